"""Host-side cost of one multi-GPU warm step (torchrun, 2+ GPUs): how long the
Python thread needs to enqueue the graph replay, the NCCL all-to-all and the
ordered adds, versus the device time of the step.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/profile_dist_host.py [--cells 100]
"""
import argparse
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    import skfem_b200 as fem
    from skfem_b200.models.poisson import laplace
    from skfem_b200.distributed import DistributedAssembler, slab_mesh_tet
    m, l2g, N, ranges = slab_mesh_tet(args.cells, args.cells, rank, world)
    for pipeline in (False, True):
        da = DistributedAssembler(laplace, fem.Basis(m, fem.ElementTetP1()), l2g, N, ranges,
                                  reuse_buffers=True, pipeline=pipeline)
        for _ in range(5):
            da.assemble()
        da.wait()
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            da.assemble()
        da.wait()
        host = (time.perf_counter() - t0) / args.steps
        ev1.record()
        torch.cuda.synchronize()
        devt = ev0.elapsed_time(ev1) / args.steps
        # parts, host side only
        ex = da.exchange
        st = da._sets[0] if pipeline else None
        out = st["out"] if pipeline else da._out
        g = st["graph"] if pipeline else da._graph
        parts = {}
        for name, fn in (("graph.replay", g.replay),
                         ("exchange.finish", lambda: ex.finish(out, zero=False))):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            parts[name] = 1e6 * (time.perf_counter() - t0) / args.steps
            torch.cuda.synchronize()
        if rank == 0:
            print("pipeline={}: device {:.1f} us/step, host enqueue {:.1f} us/step; host parts (us): {}"
                  .format(pipeline, 1e3 * devt, 1e6 * host,
                          {k: round(v, 1) for k, v in parts.items()}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
