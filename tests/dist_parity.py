"""Multi-GPU parity check against the serial oracle (test infrastructure: used by
tests/test_gpu_distributed.py and by the `parity` leg of bench.py at N > 1; the product
package never imports the oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "scikit-fem_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def parity_check(rank, world, cells_xy=12, cells_z=9):
    """Every rank's row block against the serial oracle CSR of the same global mesh, at a
    size the oracle does in about a second, through every mode of the multi-GPU path: cold
    (generic local kernels + key exchange), warm (fused local kernel writing [row block |
    send buffer] + NCCL exchange + ordered add), persistent buffers + CUDA graph, pipelined,
    and the row-range partition of one global mesh (``partition``).  indptr / indices must be
    bit-exact, values within rtol 1e-12 (+ atol 1e-12 max|A|), repeated runs bit-identical.
    Collective: call on all ranks.  Returns a dict with ``ok`` (all ranks, all modes)."""
    import torch
    import torch.distributed as dist
    from skfem_b200.basis import Basis
    from skfem_b200.distributed import DistributedAssembler, partition, slab_mesh_tet
    from skfem_b200.element import ElementTetP1
    from skfem_b200.mesh import MeshTet
    from skfem_b200.models.poisson import laplace
    res = {"modes": {}, "world": world, "cells": [cells_xy, cells_xy, cells_z * world]}
    ok = True
    max_err = 0.0
    try:
        from oracle import skfem_oracle as O
        x = np.linspace(0, 1, cells_xy + 1)
        z = np.concatenate([np.linspace(r, r + 1.0, cells_z + 1)[:-1] for r in range(world)]
                           + [[world]])
        Ag = O.assemble_bilinear(O.laplace, O.cell_basis(O.mesh_tet_tensor(x, x, z),
                                                         O.element("tet_p1")))
        scale = float(np.abs(Ag.data).max())

        def compare(A, name):
            nonlocal ok, max_err
            A.wait()
            torch.cuda.synchronize()
            blk = Ag[A.row0:A.row0 + (A.indptr.shape[0] - 1)]
            same = (np.array_equal(A.indptr.cpu().numpy(), blk.indptr) and
                    np.array_equal(A.indices.cpu().numpy(), blk.indices))
            err = float(np.abs(A.data.cpu().numpy() - blk.data).max() / scale) if same else 1.0
            good = bool(same and np.allclose(A.data.cpu().numpy(), blk.data, rtol=1e-12,
                                             atol=1e-12 * scale))
            res["modes"][name] = good
            ok = ok and good
            max_err = max(max_err, err)
            return A.data.clone()
        setups = [("slab", slab_mesh_tet(cells_xy, cells_z, rank, world)),
                  ("partition", partition(MeshTet.init_tensor(x, x, z), world, rank))]
        for tag, (m, l2g, N, ranges) in setups:
            da = DistributedAssembler(laplace, Basis(m, ElementTetP1()), l2g, N, ranges)
            compare(da.assemble(), tag + ":cold")
            warm = compare(da.assemble(), tag + ":warm")
            again = compare(da.assemble(), tag + ":warm-repeat")
            res["modes"][tag + ":deterministic"] = bool(torch.equal(warm, again))
            ok = ok and res["modes"][tag + ":deterministic"]
            for kw, name in ((dict(reuse_buffers=True), "graph"),
                             (dict(reuse_buffers=True, pipeline=True), "pipelined")):
                dp = DistributedAssembler(laplace, Basis(m, ElementTetP1()), l2g, N, ranges, **kw)
                dp.assemble()
                vals = [compare(dp.assemble(), "{}:{}".format(tag, name)) for _ in range(4)]
                dp.wait()
                torch.cuda.synchronize()
                same = all(bool(torch.equal(v, warm)) for v in vals)
                res["modes"]["{}:{}-bitwise".format(tag, name)] = same
                ok = ok and same
    except Exception as e:                      # noqa: BLE001
        ok = False
        res["error"] = "{}: {}".format(type(e).__name__, e)
    flag = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    err = torch.tensor([max_err], device="cuda", dtype=torch.float64)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    res["ok"] = bool(int(flag.item()))
    res["max_abs_err_over_max_entry"] = float(err.item())
    return res
