"""Timeline of the warm multi-GPU step (torch.profiler / CUPTI; nsys is not in the image).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/trace_dist_step.py [--cells 100] [--out gpurun_out/r2_trace_2gpu.txt]

Runs bench.py's N > 1 warm loop (pipelined DistributedAssembler on z-slabs), profiles 8 steps
on every rank and writes, for rank 0, (1) device time per kernel over the profiled steps and
(2) the device events of the last 3 steps in start order with stream, start and duration -
which shows what the step consists of beyond the two local kernels.  Written for the first and
the last rank (rank 0 of a slab decomposition only sends, the last rank only receives)."""
import argparse
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_trace_2gpu.txt"))
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import skfem_b200 as fem
    from skfem_b200.distributed import DistributedAssembler, slab_mesh_tet
    from skfem_b200.models.poisson import laplace
    m, l2g, N, ranges = slab_mesh_tet(args.cells, args.cells, rank, world)
    da = DistributedAssembler(laplace, fem.Basis(m, fem.ElementTetP1()), l2g, N, ranges,
                              reuse_buffers=True, pipeline=True)
    for _ in range(8):
        da.assemble()
    da.wait()
    dist.barrier()
    torch.cuda.synchronize()
    nsteps = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        e0.record()
        for _ in range(nsteps):
            da.assemble()
        da.wait()
        e1.record()
        torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / nsteps
    if rank in (0, world - 1):
        evs = [e for e in prof.events() if getattr(e, "device_type", None) is not None
               and str(e.device_type).endswith("CUDA")]
        evs.sort(key=lambda e: e.time_range.start)
        lines = ["# warm step of the pipelined multi-GPU path, {} GPUs, {}^3 cells per rank, "
                 "rank {} (torch.profiler, {} steps; step time under the profiler {:.4f} ms)"
                 .format(world, args.cells, rank, nsteps, step_ms), "",
                 "## device time per kernel over the profiled steps (us)"]
        agg = {}
        for e in evs:
            a = agg.setdefault(e.name[:100], [0, 0.0])
            a[0] += 1
            a[1] += e.time_range.elapsed_us()
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append("{:9.1f} us  {:4d} x  {}".format(us, n, k))
        lines += ["", "## device events of the last 3 steps in start order "
                      "(start relative to the first listed event, us)"]
        fused = [i for i, e in enumerate(evs) if "fused2" in e.name]
        start_i = fused[-3] if len(fused) >= 3 else 0
        t0 = evs[start_i].time_range.start
        for e in evs[start_i:]:
            lines.append("{:9.1f}  +{:7.1f} us  stream {:3d}  {}".format(
                e.time_range.start - t0, e.time_range.elapsed_us(),
                int(getattr(e, "stream", -1) if getattr(e, "stream", None) is not None else -1),
                e.name[:90]))
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out.replace(".txt", "_rank{}.txt".format(rank)), "w") as f:
            f.write("\n".join(lines) + "\n")
        if rank == world - 1:
            print("\n".join(lines[:40]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
