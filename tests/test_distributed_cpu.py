"""World-size-2 gloo test (CPU) of the multi-GPU host logic: row ownership,
interface-key exchange, pattern merge, ordered value reduction
(skfem_b200/distributed.py).  The per-rank local matrices come from the oracle
here; on the GPU box the same code runs on CUDA tensors over NCCL
(tests/test_gpu_distributed.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import mesh_of


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _global_mesh(cxy, cz, world):
    from oracle import skfem_oracle as O
    x = np.linspace(0, 1, cxy + 1)
    z = np.concatenate([np.linspace(r, r + 1.0, cz + 1)[:-1] for r in range(world)] + [[world]])
    return O.mesh_tet_tensor(x, x, z)


def _worker(rank, world, port, cxy, cz, out, use_partition=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import skfem_oracle as O
        from skfem_b200.distributed import (InterfaceExchange, slab_mesh_tet, balanced_ranges,
                                            partition)
        if use_partition:
            # one global mesh cut by owning row range (configs[4] style): every element once,
            # vertices renumbered ascending, contributions only to own or higher ranks' rows
            import skfem_b200 as fem
            gm = _global_mesh(cxy, cz, world)
            m, l2g, N, ranges = partition(fem.MeshTet(gm.p, gm.t), world, rank)
            counts = [partition(fem.MeshTet(gm.p, gm.t), world, r)[0].nelements
                      for r in range(world)]
            assert sum(counts) == gm.t.shape[1] and max(counts) - min(counts) <= 0.2 * max(counts)
            assert np.all(np.diff(l2g) > 0) and ranges[0] == 0 and ranges[-1] == gm.p.shape[1]
            assert l2g.min() >= ranges[rank]
        else:
            m, l2g, N, ranges = slab_mesh_tet(cxy, cz, rank, world)
        Aloc = O.assemble_bilinear(O.laplace, O.cell_basis(mesh_of(dict(p=m.p, t=m.t), "tet"),
                                                           O.element("tet_p1")))
        lrow = np.repeat(np.arange(Aloc.shape[0]), np.diff(Aloc.indptr))
        g = torch.as_tensor(l2g)
        ex = InterfaceExchange(g[torch.as_tensor(lrow)], g[torch.as_tensor(Aloc.indices).long()],
                               ranges, N)
        d1 = ex.reduce(torch.as_tensor(Aloc.data))
        d2 = ex.reduce(torch.as_tensor(Aloc.data))
        assert torch.equal(d1, d2)                                   # deterministic
        # direct-write layout used by the fused kernel: [row block | send buffer]
        out_ = torch.full((ex.nnz + ex.nsend,), float("nan"), dtype=torch.float64)
        out_[ex.slot_map] = torch.as_tensor(Aloc.data)
        d3 = ex.finish(out_)
        np.testing.assert_allclose(d3.numpy(), d1.numpy(), rtol=1e-15, atol=0)
        Ag = O.assemble_bilinear(O.laplace, O.cell_basis(_global_mesh(cxy, cz, world),
                                                         O.element("tet_p1")))
        blk = Ag[ex.row0:ex.row0 + ex.nrows]
        assert np.array_equal(ex.indptr.numpy(), blk.indptr)
        assert np.array_equal(ex.indices.numpy(), blk.indices)
        np.testing.assert_allclose(d1.numpy(), blk.data, rtol=1e-12,
                                   atol=1e-12 * np.abs(blk.data).max())
        assert ex.bytes_per_exchange > 0 or rank == world - 1   # the top part only receives
        assert list(balanced_ranges(10, 3)) == [0, 4, 7, 10]
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_interface_exchange_gloo(world):
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, 5, 3, out), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))


@pytest.mark.parametrize("world", [2, 3])
def test_row_range_partition_gloo(world):
    """distributed.partition (BASELINE configs[4]): the row blocks assembled from the parts of
    one global mesh equal the serial oracle CSR."""
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, 5, 3, out, True), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))
