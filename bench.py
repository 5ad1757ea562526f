#!/usr/bin/env python
"""Benchmark of the assembly hot path (BASELINE.json metric: assembly elements/s
and CSR nnz/s, FP64, P1-tet Laplace).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C]

Workload at N=1: BASELINE.json configs[1] - MeshTet.init_tensor with 101 points
per side (6 000 000 P1 tetrahedra, 1 030 301 DOFs, 7 150 901 CSR nnz), Laplace.
A "step" is one warm re-assembly (sparsity plan cached) of the whole mesh into
CSR values with p, t and the plan resident in HBM.  N>1 (torchrun, one rank per
GPU): weak scaling - every rank owns one z-slab of C*C*C cells of a mesh N
slabs tall (multi-GPU).

One JSON line on stdout (rank 0).  Keys follow the driver contract plus
`roofline`, `cpu_baseline`, `e2e`, `gpu_launches`, `clocks`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "scikit-fem_b200"))
sys.path.insert(0, ROOT)

METRIC = "assembly elements/s (FP64 P1-tet Laplace -> CSR)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------
def cpu_assemble(cells, nthreads):
    """One reference-style cold assembly (Basis + laplace.assemble) of an
    init_tensor mesh with `cells` cells per side; returns (seconds, nel, nnz)."""
    from oracle import skfem_oracle as O
    x = np.linspace(0, 1, cells + 1)
    m = O.mesh_tet_tensor(x, x, x)
    t0 = time.perf_counter()
    b = O.cell_basis(m, O.element("tet_p1"))
    idx, data, shape = O.bilinear_coo(O.laplace, b, nthreads=nthreads)
    A = O.coo_to_csr(idx, data, shape)
    dt = time.perf_counter() - t0
    return dt, m.t.shape[1], A.nnz


def run_reference(args):
    cores = os.cpu_count() or 1
    nthreads = min(cores, 16)
    cells = args.ref_cells
    for _ in range(args.warmup):
        cpu_assemble(min(cells, 20), nthreads)
    times = []
    for _ in range(args.steps):
        dt, nel, nnz = cpu_assemble(cells, nthreads)
        times.append(dt)
    total = sum(times)
    val = nel * args.steps / total
    sample = ("oracle port (numpy/scipy restatement of Basis + BilinearForm._assemble + "
              "eliminate_zeros/tocsr), init_tensor {}^3 cells = {} P1 tets per step, "
              "nthreads={} over the Nbfun^2 loop".format(cells, nel, nthreads))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "nnz_per_s": nnz * args.steps / total,
        "config": {"workload": "MeshTet.init_tensor 101^3 pts ElementTetP1 laplace (configs[1]); "
                               "CPU arm timed on a bounded sample of the same mesh family",
                   "sample_cells_per_side": cells, "elements": nel},
        "cpu_baseline": {"value": val, "unit": "elements/s", "cores": nthreads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import skfem_b200 as fem
    from skfem_b200 import _lib
    from skfem_b200.models.poisson import laplace

    from skfem_b200 import form as _form
    _form.set_options(fused=not args.no_fused, fused_tile=args.tile, fused_threads=args.threads,
                      fused_ring=args.ring, fused_arith=args.arith,
                      fused_spread=not args.no_spread, fused_l2_persist=args.l2_persist,
                      fused_renumber=not args.no_renumber, fused_tiling=args.tiling,
                      fused_version=args.fused_version, fused2_tile=args.tile2,
                      fused2_ring=args.ring2, fused2_pool=args.pool, fused2_ctas=args.ctas,
                      fused2_S=args.super_tiles)
    cells = args.cells
    da = None
    if world == 1:
        x = np.linspace(0, 1, cells + 1)
        m = fem.MeshTet.init_tensor(x, x, x)
    else:
        # weak scaling: rank r assembles the z-slab [r, r+1] (cells^3 cells) of a
        # mesh `world` slabs tall; interface rows go to their owner over NCCL
        from skfem_b200.distributed import DistributedAssembler, slab_mesh_tet
        m, l2g, Nglob, ranges = slab_mesh_tet(cells, cells, rank, world)
    basis = fem.Basis(m, fem.ElementTetP1())
    nel = m.nelements

    # cold assembly: builds and caches the plan (not part of the warm step)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world == 1:
        A = laplace.assemble_device(basis)
    else:
        # pipelined: the interface exchange of step i overlaps the local kernels of
        # step i+1 (two output buffer sets); SKB_PIPELINE=0 serialises them again
        da = DistributedAssembler(laplace, basis, l2g, Nglob, ranges, reuse_buffers=True,
                                  graph_exchange=os.environ.get("SKB_GRAPH_EXCHANGE") == "1",
                                  pipeline=os.environ.get("SKB_PIPELINE", "1") == "1",
                                  sm_reserve=int(os.environ.get("SKB_SM_RESERVE", "0")))
        A = da.assemble()
    torch.cuda.synchronize()
    cold_ms = 1e3 * (time.perf_counter() - t0)
    nnz = A.nnz
    if world > 1:
        tn = torch.tensor([nnz], device="cuda", dtype=torch.int64)
        dist.all_reduce(tn)
        nnz_total = int(tn.item())
    else:
        nnz_total = nnz

    if args.debug_flags:
        _lib.lib().skb_debug_flags(args.debug_flags)   # profiling only: results invalid

    if world == 1:
        out = torch.empty(nnz, dtype=torch.float64, device="cuda")

        def step_eager():
            return laplace.assemble_device(basis, out=out)
    else:
        step_eager = da.assemble     # local fused assembly + interface exchange + ordered add

    for _ in range(2):
        step_eager()                 # builds the fused tile plan on the first warm call
    torch.cuda.synchronize()
    launches_per_step = None
    graph = None
    if not args.no_graph and world == 1:
        # the warm step is launch bound from Python: capture it once, replay it
        _lib.lib().skb_launch_count(1)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_eager()
        launches_per_step = int(_lib.lib().skb_launch_count(1))
        step = graph.replay
    else:
        step = step_eager
    for _ in range(max(args.warmup, 3)):
        step()
    # the clock sampler (an nvidia-smi subprocess on rank 0) is started BEFORE the
    # barrier: spawning it takes ~1 ms, which the other ranks would otherwise spend
    # inside their timed region waiting for rank 0 at the first exchange
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        da.wait()
        torch.cuda.synchronize()
        dist.barrier()
    torch.cuda.synchronize()
    _lib.lib().skb_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    if da is not None:
        da.wait()                    # the last exchanges (side stream) are part of the K steps
    ev1.record()
    torch.cuda.synchronize()
    launches = int(_lib.lib().skb_launch_count(0))
    if graph is not None:            # replays launch the captured kernels, not the C API
        launches = launches_per_step * args.steps
    elif da is not None:             # the assembler replays its own graph: fused + combine
        launches = 2 * args.steps
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(tms.item())
    ms_step = ms / args.steps
    value = nel * world / (ms_step * 1e-3)

    # side measurement (not the headline): the same warm step with the fused kernel's
    # opt-in fast arithmetic (FMA + one reciprocal per element, values within rtol 1e-12)
    fast_line = None
    if world == 1 and graph is not None and args.arith == "exact" and not args.no_fused:
        _form.set_options(fused_arith="fast")
        try:
            step_eager()
            torch.cuda.synchronize()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                step_eager()
            for _ in range(3):
                g2.replay()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.steps):
                g2.replay()
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1) / args.steps
            fast_line = {"ms_per_step": fms, "value": nel / (fms * 1e-3), "unit": "elements/s",
                         "note": "opt-in set_options(fused_arith='fast'); not the headline"}
        finally:
            _form.set_options(fused_arith="exact")

    # ---- end to end: host buffers in, scipy CSR out, through the public API ----
    e2e = None
    if not args.no_e2e:
        p_pin = torch.from_numpy(m.p).pin_memory()
        t_pin = torch.from_numpy(m.t).pin_memory()
        p_host, t_host = p_pin.numpy(), t_pin.numpy()

        def e2e_step():
            mm = fem.MeshTet(p_host, t_host)
            bb = fem.Basis(mm, fem.ElementTetP1())
            if world == 1:
                return laplace.assemble(bb)
            # every rank: its slab in, its row block of the global matrix out
            return DistributedAssembler(laplace, bb, l2g, Nglob, ranges).assemble() \
                .to_scipy_block()
        # warm-up: the first calls also populate torch's pinned-host allocator pool for the
        # three result arrays (cudaHostAlloc of 90 MB costs tens of ms; tools/profile_e2e.py)
        for _ in range(max(args.warmup, 4)):
            Ah = e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        k_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        trace = []
        for _ in range(k_e2e):
            t1 = time.perf_counter()
            Ah = e2e_step()
            trace.append((1e3 * (time.perf_counter() - t1),
                          torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k_e2e
        if os.environ.get("SKB_BENCH_TRACE"):
            print("e2e per-iteration (ms, MiB allocated, MiB reserved):", trace, file=sys.stderr)
        if world > 1:
            tdt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        e2e = {"value": nel * world / dt, "unit": "elements/s",
               "h2d_bytes_per_step": int(m.p.nbytes + m.t.nbytes),
               "d2h_bytes_per_step": int(Ah.data.nbytes + Ah.indices.nbytes + Ah.indptr.nbytes),
               "ms_per_step": 1e3 * dt,
               "what": "cold: MeshTet(p,t) + Basis + laplace.assemble -> scipy csr_matrix, "
                       "plan build included, pinned host inputs"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    nverts = m.p.shape[1]
    fused_stats = None
    for k, v in basis._plans.items():
        if isinstance(k, tuple) and k and k[0] in ("fused", "fused-mapped") and v is not None:
            if getattr(v, "version", 1) == 2:
                from skfem_b200 import fused2 as _fused2
                fused_stats = _fused2.stats(v)
            else:
                from skfem_b200 import fused as _fused
                fused_stats = _fused.stats(v)
                fused_stats["tile"], fused_stats["threads"], fused_stats["ring"] = \
                    v.T, v.threads, v.ring
    algo_bytes = 4 * 4 * nel + 8 * 3 * nverts + 8 * nnz   # t + p + CSR data (warm)
    achieved = algo_bytes / (ms_step * 1e-3) / 1e9
    # DRAM traffic per step from the committed `ncu --set full` captures (same workload/path)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        if (tr["cells"] == cells and tr["path"] == ("fused" if fused_stats else "generic")
                and world == 1):
            traffic = tr["dram_bytes_per_step"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "nnz_per_s": nnz_total / (ms_step * 1e-3),
        "config": {"workload": ("MeshTet.init_tensor {}^3 pts ElementTetP1 laplace "
                                "(BASELINE configs[1]) warm re-assembly into CSR".format(cells + 1)
                                if world == 1 else
                                "MeshTet.init_tensor z-slab of {0}^3 cells per GPU, {1} slabs, "
                                "ElementTetP1 laplace, warm re-assembly into a row-partitioned "
                                "CSR, interface rows exchanged over NCCL{2}".format(
                                    cells, world, " (exchange of step i overlapped with the "
                                    "local kernels of step i+1)" if da.pipeline else "")),
                   "elements_per_gpu": nel, "dofs_per_gpu": basis.N, "nnz_per_gpu": nnz,
                   "nnz_total": nnz_total,
                   "interface_bytes_sent_rank0": (da.exchange.bytes_per_exchange
                                                  if da is not None else 0),
                   "l2": "no flush: per-step working set (t, local data, plan) exceeds the 126 MB L2",
                   "cold_plan_build_ms": cold_ms,
                   "path": "fused" if fused_stats else "generic", "fused_plan": fused_stats},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": algo_bytes,
                     "note": "whole warm step (all kernels of the step) vs compulsory bytes "
                             "t + p + CSR data"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if fast_line is not None:
        line["fast_arith"] = fast_line
    if world == 1 and not args.no_cpu:
        cores = 1
        dt, cnel, cnnz = cpu_assemble(args.ref_cells, 0)
        line["cpu_baseline"] = {
            "value": cnel / dt, "unit": "elements/s", "cores": cores, "kind": "port",
            "sample": "oracle port, single thread (reference default nthreads=0), init_tensor "
                      "{}^3 cells = {} P1 tets, Basis + assemble + tocsr, {:.1f} s".format(
                          args.ref_cells, cnel, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100, help="cells per side per GPU")
    ap.add_argument("--ref-cells", type=int, default=60, dest="ref_cells",
                    help="cells per side of the CPU sample (60 -> 1.3 M tets)")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e")
    ap.add_argument("--debug-flags", type=int, default=0, dest="debug_flags",
                    help="profiling aid (skb_debug_flags): 1 skip P1, 2 skip P2; invalid results")
    ap.add_argument("--no-graph", action="store_true", dest="no_graph",
                    help="launch the warm step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-fused", action="store_true", dest="no_fused",
                    help="time the generic two-kernel path instead of the fused P1 kernel")
    ap.add_argument("--arith", default="exact", choices=["exact", "fast"],
                    help="fused-kernel arithmetic: exact = reference operation order (default, "
                         "the headline), fast = FMA + reciprocal (values within rtol 1e-12)")
    ap.add_argument("--no-renumber", action="store_true", dest="no_renumber",
                    help="keep the tile-local vertex ids in global-id order (no bank colouring)")
    ap.add_argument("--no-spread", action="store_true", dest="no_spread",
                    help="keep the COO order inside the P2 lists (no plan-time bank spreading)")
    ap.add_argument("--no-l2-persist", action="store_false", dest="l2_persist",
                    help="do not pin the fused path's scratch array in L2 between its two kernels")
    ap.add_argument("--tiling", default="morton", choices=["morton", "kd"],
                    help="fused plan: how elements are cut into tiles (kd = compact k-d boxes, "
                         "opt-in until timed)")
    ap.add_argument("--fused-version", type=int, default=2, dest="fused_version",
                    help="2: super-tile / pool kernel (default), 1: first-generation kernel")
    ap.add_argument("--tile2", type=int, default=256, help="v2: elements per tile = threads per CTA")
    ap.add_argument("--ring2", type=int, default=3, help="v2: record buffers per CTA")
    ap.add_argument("--pool", type=int, default=2048, help="v2: accumulators per super-tile")
    ap.add_argument("--ctas", type=int, default=0, help="v2: CTAs per SM (0 = as many as fit)")
    ap.add_argument("--super-tiles", type=int, default=None, dest="super_tiles",
                    help="v2: tiles per super-tile (default: as many as the pool allows)")
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--threads", type=int, default=480,
                    help="reduce threads per CTA of the fused kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference(args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
