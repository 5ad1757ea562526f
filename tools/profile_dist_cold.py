"""Phase timing of the cold multi-GPU call (what bench.py's e2e measures at N > 1):
local cold assembly, exchange plan (key exchange, sorted unique, slot maps), value
reduction, row block to host.  Runs under torchrun with any world size, also 1.

    python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 \
        tools/profile_dist_cold.py [--cells 100]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    import skfem_b200 as fem
    from skfem_b200.models.poisson import laplace
    from skfem_b200.distributed import (DistributedAssembler, DistributedCSR, InterfaceExchange,
                                        slab_mesh_tet)
    m, l2g, N, ranges = slab_mesh_tet(args.cells, args.cells, rank, world)
    p = torch.from_numpy(m.p).pin_memory().numpy()     # pinned host inputs, like bench.py
    t = torch.from_numpy(m.t).pin_memory().numpy()

    def T(label, fn, n=4):
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize()
        if rank == 0:
            print("{:46s} {:8.2f} ms".format(label, 1e3 * (time.perf_counter() - t0) / n), flush=True)
        return r

    def full():
        bb = fem.Basis(fem.MeshTet(p, t), fem.ElementTetP1())
        return DistributedAssembler(laplace, bb, l2g, N, ranges).assemble().to_scipy_block()
    T("full cold call (bench e2e at this N)", full)
    A = T("  Basis + local cold assemble_device", lambda: laplace.assemble_device(
        fem.Basis(fem.MeshTet(p, t), fem.ElementTetP1())))
    dev = A.data.device
    l2g_d = torch.as_tensor(np.asarray(l2g, dtype=np.int64), device=dev)
    counts = (A.indptr[1:] - A.indptr[:-1]).long()
    lrow = torch.repeat_interleave(torch.arange(A.shape[0], device=dev), counts)
    grow, gcol = l2g_d[lrow], l2g_d[A.indices.long()]
    ex = T("  InterfaceExchange(...) [plan]", lambda: InterfaceExchange(grow, gcol, ranges, N))
    data = T("  exchange.reduce(values)", lambda: ex.reduce(A.data))
    blk = DistributedCSR(ex.indptr, ex.indices, data, ex.row0, (N, N))
    T("  row block to host (to_scipy_block)", blk.to_scipy_block)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
