"""Two-rank NCCL test of the multi-GPU path (needs >= 2 GPUs; skipped otherwise):
every rank assembles its z-slab with the fused kernel, interface rows are
exchanged, and each rank's row block must match the serial oracle CSR."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cxy, cz, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import skfem_b200 as fem
        from skfem_b200.models.poisson import laplace
        from skfem_b200.distributed import DistributedAssembler, slab_mesh_tet
        from oracle import skfem_oracle as O
        m, l2g, N, ranges = slab_mesh_tet(cxy, cz, rank, world)
        da = DistributedAssembler(laplace, fem.Basis(m, fem.ElementTetP1()), l2g, N, ranges)
        A1 = da.assemble()           # cold (generic local path)
        A2 = da.assemble()           # warm (fused local path)
        A3 = da.assemble()
        assert torch.equal(A2.data, A3.data)
        x = np.linspace(0, 1, cxy + 1)
        z = np.concatenate([np.linspace(r, r + 1.0, cz + 1)[:-1] for r in range(world)]
                           + [[world]])
        Ag = O.assemble_bilinear(O.laplace, O.cell_basis(O.mesh_tet_tensor(x, x, z),
                                                         O.element("tet_p1")))
        blk = Ag[A2.row0:A2.row0 + (A2.indptr.shape[0] - 1)]
        for A in (A1, A2):
            assert np.array_equal(A.indptr.cpu().numpy(), blk.indptr)
            assert np.array_equal(A.indices.cpu().numpy(), blk.indices)
            np.testing.assert_allclose(A.data.cpu().numpy(), blk.data, rtol=1e-12,
                                       atol=1e-12 * np.abs(blk.data).max())
        # re-assembly loop modes: persistent buffers + CUDA graph, and the pipelined
        # variant (exchange of step i overlapped with the kernels of step i+1)
        for kw in (dict(reuse_buffers=True), dict(reuse_buffers=True, pipeline=True)):
            dp = DistributedAssembler(laplace, fem.Basis(m, fem.ElementTetP1()), l2g, N, ranges,
                                      **kw)
            dp.assemble()
            blocks = []
            for _ in range(5):           # back to back, no host synchronisation in between
                blocks.append(dp.assemble().wait().data.clone())
            dp.wait()
            torch.cuda.synchronize()
            for d in blocks:
                assert torch.equal(d, A2.data), kw
            last = [dp.assemble() for _ in range(6)][-2:]     # fully overlapped
            dp.wait()
            torch.cuda.synchronize()
            for A in last:
                assert torch.equal(A.data, A2.data), kw
        # the same check bench.py prints as `parity` at N > 1 (all modes, slab + partition)
        from dist_parity import parity_check
        par = parity_check(rank, world)
        assert par["ok"], par
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_matches_oracle():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 12, 9, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
