#!/usr/bin/env python
"""Benchmark of the assembly hot path (BASELINE.json metric: assembly elements/s
and CSR nnz/s, FP64, P1-tet Laplace).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C]

Workload at N=1: BASELINE.json configs[1] - MeshTet.init_tensor with 101 points
per side (6 000 000 P1 tetrahedra, 1 030 301 DOFs, 7 150 901 CSR nnz), Laplace.
A "step" is one warm re-assembly (sparsity plan cached) of the whole mesh into
CSR values with t and the plan resident in HBM and *new vertex coordinates every
step*: the steps alternate between two coordinate arrays (the grid and an
anisotropically scaled copy, both resident), so consecutive steps produce
different matrices and every step re-derives the zero mask of every local matrix
against the cached value-dependent pattern (basis.update_points).  N>1
(torchrun, one rank per GPU): weak scaling - every rank owns one z-slab of C*C*C
cells of a mesh N slabs tall (multi-GPU); `--strong` partitions one
global mesh instead (configs[4]).

One JSON line on stdout (rank 0).  Keys follow the driver contract plus
`roofline`, `cpu_baseline`, `e2e`, `gpu_launches`, `clocks`, `checks`, `parity`.
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
SCALE_B = (1.25, 0.8, 1.1)     # the second coordinate set: anisotropic scaling of the grid


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(cells, world, strong=False, pipeline=False):
    if world == 1:
        return ("MeshTet.init_tensor {}^3 pts ElementTetP1 laplace (BASELINE configs[1]) warm "
                "re-assembly into CSR, vertex coordinates changing every step".format(cells + 1))
    if strong:
        return ("MeshTet.init_tensor {0}^3 pts ElementTetP1 laplace (BASELINE configs[4]) "
                "element-partitioned over {1} GPUs by owning row range, warm re-assembly into a "
                "row-partitioned CSR, interface rows exchanged over NCCL".format(cells + 1, world))
    return ("MeshTet.init_tensor z-slab of {0}^3 cells per GPU, {1} slabs, ElementTetP1 laplace, "
            "warm re-assembly into a row-partitioned CSR, interface rows exchanged over NCCL{2}"
            .format(cells, world, " (exchange of step i overlapped with the local kernels of "
                    "step i+1)" if pipeline else ""))


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
# reference arm / cpu baseline: the reference itself (oracle/_ref, installed by
# tools/install_ref.sh) or, where it is absent, the oracle port - on host cores
# --------------------------------------------------------------------------
def reference_package():
    """The unmodified scikit-fem from oracle/_ref, or None."""
    path = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(path, "skfem")):
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    try:
        import skfem
        return skfem
    except Exception:
        return None


def cpu_assemble(cells, nthreads):
    """One reference cold assembly - Basis(m, ElementTetP1()) + laplace.assemble(basis), the
    timed part of docs/examples/performance.py:19-24 without the load vector - of an
    init_tensor mesh with `cells` cells per side; returns (seconds, nel, nnz, kind)."""
    skfem = reference_package()
    x = np.linspace(0, 1, cells + 1)
    if skfem is not None:
        from skfem.models.poisson import laplace
        form = laplace if not nthreads else skfem.BilinearForm(laplace.form, nthreads=nthreads)
        m = skfem.MeshTet.init_tensor(x, x, x)
        t0 = time.perf_counter()
        A = form.assemble(skfem.Basis(m, skfem.ElementTetP1()))
        dt = time.perf_counter() - t0
        return dt, m.nelements, A.nnz, "reference"
    from oracle import skfem_oracle as O
    m = O.mesh_tet_tensor(x, x, x)
    t0 = time.perf_counter()
    b = O.cell_basis(m, O.element("tet_p1"))
    idx, data, shape = O.bilinear_coo(O.laplace, b, nthreads=nthreads)
    A = O.coo_to_csr(idx, data, shape)
    dt = time.perf_counter() - t0
    return dt, m.t.shape[1], A.nnz, "port"


def run_reference(args):
    cores = os.cpu_count() or 1
    nthreads = min(cores, 16)
    for _ in range(args.warmup):
        cpu_assemble(20, nthreads)
    # bounded sample: the largest mesh of the same family whose K steps fit the time budget,
    # from the rate measured on a 40^3-cell calibration run; the full configs[1] size if it fits
    dt, nel, _, kind = cpu_assemble(40, nthreads)
    rate = nel / dt
    cells = int(min(args.cells,
                    max(20, (rate * args.ref_seconds / args.steps / 6.0) ** (1.0 / 3.0))))
    if args.ref_cells:
        cells = args.ref_cells
    times = []
    for _ in range(args.steps):
        dt, nel, nnz, kind = cpu_assemble(cells, nthreads)
        times.append(dt)
    total = sum(times)
    val = nel * args.steps / total
    what = ("the unmodified reference (scikit-fem 12.0.1 from oracle/_ref)" if kind == "reference"
            else "oracle port (numpy/scipy restatement of the reference's path)")
    share = "all" if cells == args.cells else "{:.0%}".format(nel / (6.0 * args.cells ** 3))
    sample = ("{}: Basis + laplace.assemble -> csr_matrix on init_tensor {}^3 cells = {} P1 tets "
              "per step ({} of the configs[1] mesh), BilinearForm(nthreads={}) over the Nbfun^2 "
              "loop, {} host cores".format(what, cells, nel, share, nthreads, cores))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if (world > 1 and args.strong) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "nnz_per_s": nnz * args.steps / total,
        "config": {"workload": workload_name(args.cells, world, args.strong,
                                             os.environ.get("SKB_PIPELINE", "1") == "1")},
        "cpu_baseline": {"value": val, "unit": "elements/s", "cores": nthreads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------
# DRAM traffic of the warm step, measured live: a child run of this script under ncu
# --------------------------------------------------------------------------
def traffic_child(args):
    """Minimal warm loop for the ncu child: cold assembly, then warm steps."""
    import torch
    import skfem_b200 as fem
    from skfem_b200.models.poisson import laplace
    apply_options(args)
    x = np.linspace(0, 1, args.cells + 1)
    basis = fem.Basis(fem.MeshTet.init_tensor(x, x, x), fem.ElementTetP1())
    A = laplace.assemble_device(basis)
    out = torch.empty(A.nnz, dtype=torch.float64, device="cuda")
    for _ in range(4):
        laplace.assemble_device(basis, out=out)
    torch.cuda.synchronize()


def measure_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of one warm step (all its kernels), from
    `ncu` on a child process running the same configuration.  (None, reason) if unavailable."""
    import csv
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = os.path.join(ROOT, "gpurun_out", "traffic_ncu.csv")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    pat = "regex:fused|combine" if not args.no_fused else "regex:local_affine|csr_reduce"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control",
           "none", "-k", pat, "--csv", "--log-file", log, sys.executable,
           os.path.abspath(__file__), "--traffic-child", "--cells", str(args.cells)]
    for flag in ("fused_version", "tile2", "ring2", "pool", "ctas", "ept", "tile", "ring",
                 "threads"):
        cmd += ["--" + flag.replace("_", "-"), str(getattr(args, flag))]
    if args.super_tiles:
        cmd += ["--super-tiles", str(args.super_tiles)]
    if args.no_fused:
        cmd.append("--no-fused")
    try:
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300,
                       check=True)
        with open(log) as f:
            lines = [ln for ln in f if ln.startswith('"')]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per_launch = {}
        for r in csv.DictReader(lines):
            val = float(r["Metric Value"].replace(",", "")) * unit.get(r["Metric Unit"], 1.0)
            per_launch.setdefault(int(r["ID"]), [r["Kernel Name"], 0.0])[1] += val
        if not per_launch:
            return None, "no kernels captured"
        ids = sorted(per_launch)
        # the last warm step = the trailing launches back to the last main-kernel launch
        main = [i for i in ids if "combine" not in per_launch[i][0]
                and "reduce" not in per_launch[i][0]]
        last = [i for i in ids if i >= main[-1]]
        detail = {}
        for i in last:
            name = per_launch[i][0].split("(")[0].split("<")[0].split("::")[-1].strip()
            detail[name] = detail.get(name, 0.0) + per_launch[i][1]
        return sum(detail.values()), detail
    except Exception as e:                      # noqa: BLE001
        return None, "ncu run failed: {}".format(e)


def apply_options(args):
    from skfem_b200 import form as _form
    _form.set_options(fused=not args.no_fused, fused_tile=args.tile, fused_threads=args.threads,
                      fused_ring=args.ring, fused_arith=args.arith,
                      fused_spread=not args.no_spread, fused_l2_persist=args.l2_persist,
                      fused_renumber=not args.no_renumber, fused_tiling=args.tiling,
                      fused_version=args.fused_version, fused2_tile=args.tile2,
                      fused2_ring=args.ring2, fused2_pool=args.pool, fused2_ctas=args.ctas,
                      fused2_S=args.super_tiles, fused2_ept=args.ept,
                      hex_sumfact=not args.no_hex_sumfact,
                      element_major=not args.no_element_major)
    return _form


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

    _form = apply_options(args)
    cells = args.cells
    da = None
    parity = None
    if world == 1:
        x = np.linspace(0, 1, cells + 1)
        m = fem.MeshTet.init_tensor(x, x, x)
    else:
        from skfem_b200.distributed import DistributedAssembler, slab_mesh_tet, partition
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from dist_parity import parity_check       # checker only (uses the oracle)
        # NCCL parity against the oracle at a size the oracle does in a second: every rank's
        # row block, eager / persistent-buffer / pipelined modes (printed as `parity`)
        if not args.no_check:
            parity = parity_check(rank, world)
        if args.strong:
            # configs[4]: one global mesh, elements assigned to the rank owning the row of
            # their smallest DOF (contiguous, element-balanced row ranges)
            x = np.linspace(0, 1, cells + 1)
            m, l2g, Nglob, ranges = partition(fem.MeshTet.init_tensor(x, x, x), world, rank)
        else:
            # weak scaling: rank r assembles the z-slab [r, r+1] (cells^3 cells) of a mesh
            # `world` slabs tall; interface rows go to their owner over NCCL
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
                                  sm_reserve=int(os.environ.get("SKB_SM_RESERVE", "0")),
                                  depth=int(os.environ.get("SKB_PIPE_DEPTH", "2")))
        A = da.assemble()
    torch.cuda.synchronize()
    cold_ms = 1e3 * (time.perf_counter() - t0)
    nnz = A.nnz
    if world > 1:
        tn = torch.tensor([nnz, nel], device="cuda", dtype=torch.int64)
        dist.all_reduce(tn)
        nnz_total, nel_total = int(tn[0].item()), int(tn[1].item())
    else:
        nnz_total, nel_total = nnz, nel

    if args.debug_flags:
        _lib.lib().skb_debug_flags(args.debug_flags)   # profiling only: results invalid

    checks = None
    graphs = None
    launches_per_step = None
    moving = False
    if world == 1:
        out = torch.empty(nnz, dtype=torch.float64, device="cuda")
        p_a = basis._dev()["p"]
        p_b = (p_a * torch.tensor(SCALE_B, device="cuda", dtype=torch.float64)[:, None]
               ).contiguous()
        moving = not args.static_points and not args.no_fused and args.fused_version == 2

        def assemble_with(p):
            if moving:
                basis.update_points(p, host=False, adopt=True)
            return laplace.assemble_device(basis, out=out)
        laplace.assemble_device(basis, out=out)      # first warm call: plan of the fused path
        for p in (p_b, p_a):
            assemble_with(p)                         # moved mesh: pattern re-validated
        torch.cuda.synchronize()
        state = {"i": 0}
        if not args.no_graph:
            # the warm step is launch bound from Python: one captured graph per coordinate
            # buffer (the pointer is baked into the graph), replayed alternately
            graphs = []
            for p in ((p_a, p_b) if moving else (p_a,)):
                assemble_with(p)
                torch.cuda.synchronize()
                _lib.lib().skb_launch_count(1)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    laplace.assemble_device(basis, out=out)
                launches_per_step = int(_lib.lib().skb_launch_count(1))
                graphs.append(g)

            def step():
                graphs[state["i"] % len(graphs)].replay()
                state["i"] += 1
        else:
            def step():
                assemble_with((p_a, p_b)[state["i"] % 2])
                state["i"] += 1
    else:
        step = da.assemble     # local fused assembly + interface exchange + ordered add

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
    if graphs is not None:           # replays launch the captured kernels, not the C API
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
    value = nel_total / (ms_step * 1e-3)

    if world == 1 and not args.debug_flags:
        # the timed steps really assembled two different matrices, and the right ones: exact
        # energy of a linear field on both geometries, pattern still valid
        checks = {}
        u0 = np.array([2.0, -3.0, 0.5])
        Acsr = None
        for name, p, s in (("grid", p_a, (1.0, 1.0, 1.0)), ("scaled", p_b, SCALE_B)):
            if name == "scaled" and not moving:
                continue
            assemble_with(p)
            u = torch.as_tensor(u0, device="cuda") @ p
            Acsr = torch.sparse_csr_tensor(A.indptr.long(), A.indices.long(), out, size=A.shape)
            energy = float(u @ (Acsr @ u))
            # |grad u|^2 * volume: u = a.x on the scaled box has gradient a (unchanged)
            exact = float((u0 ** 2).sum() * np.prod(s))
            checks["energy_rel_err_" + name] = abs(energy - exact) / exact
            checks["checksum_" + name] = float(out.abs().sum())
        for k, v in basis._plans.items():
            if isinstance(k, tuple) and k and k[0] == "fused" and hasattr(v, "flag"):
                checks["pattern_flag"] = int(v.flag.item())
        errs = [v for k, v in checks.items() if k.startswith("energy_rel_err")]
        checks["ok"] = bool(max(errs) < 1e-10 and checks.get("pattern_flag", 0) == 0 and
                            (not moving or checks["checksum_grid"] != checks["checksum_scaled"]))
        checks["moving_points"] = bool(moving)
        if moving:
            basis.update_points(p_a, host=False, adopt=True)
            laplace.assemble_device(basis, out=out)

    # side measurement (not the headline): the same warm step with the fused kernel's
    # opt-in fast arithmetic (FMA + one reciprocal per element, values within rtol 1e-12)
    fast_line = None
    if world == 1 and graphs is not None and args.arith == "exact" and not args.no_fused \
            and not args.debug_flags:
        _form.set_options(fused_arith="fast")
        try:
            laplace.assemble_device(basis, out=out)
            torch.cuda.synchronize()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                laplace.assemble_device(basis, out=out)
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
            # every rank: its part in, its row block of the global matrix out
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
        for _ in range(k_e2e):
            Ah = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k_e2e
        if world > 1:
            tdt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        e2e = {"value": nel_total / dt, "unit": "elements/s",
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
    traffic, traffic_detail = None, None
    if world == 1 and not args.no_traffic and not args.debug_flags:
        traffic, traffic_detail = measure_traffic(args)
    line = {
        "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong" if (world > 1 and args.strong) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "nnz_per_s": nnz_total / (ms_step * 1e-3),
        "config": {"workload": workload_name(cells, world, args.strong,
                                             da is not None and da.pipeline),
                   "elements_per_gpu": nel, "elements_total": nel_total,
                   "dofs_per_gpu": basis.N, "nnz_per_gpu": nnz, "nnz_total": nnz_total,
                   "interface_bytes_sent_rank0": (da.exchange.bytes_per_exchange
                                                  if da is not None else 0),
                   "l2": "no flush: per-step working set (tile records, coordinates, CSR values, "
                         "tile partials) exceeds the 126 MB L2",
                   "cold_plan_build_ms": cold_ms,
                   "path": "fused" if fused_stats else "generic", "fused_plan": fused_stats},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": algo_bytes,
                     "traffic_over_algorithmic": (traffic / algo_bytes) if traffic else None,
                     "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum of the "
                                        "kernels of one warm step, child run of this command"
                                        if traffic else traffic_detail),
                     "traffic_by_kernel": traffic_detail if traffic else None,
                     "note": "whole warm step (all kernels of the step) vs compulsory bytes "
                             "t + p + CSR data"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "checks": checks,
    }
    if parity is not None:
        line["parity"] = parity
        line["parity_ok"] = bool(parity.get("ok"))
    if fast_line is not None:
        line["fast_arith"] = fast_line
    if world == 1 and not args.no_cpu:
        dt, cnel, cnnz, kind = cpu_assemble(args.cpu_cells, 0)
        line["cpu_baseline"] = {
            "value": cnel / dt, "unit": "elements/s", "cores": 1, "kind": kind,
            "sample": "{}, single thread (reference default nthreads=0), init_tensor {}^3 cells "
                      "= {} P1 tets, Basis + laplace.assemble -> csr_matrix, {:.1f} s".format(
                          "the unmodified reference (oracle/_ref)" if kind == "reference"
                          else "oracle port", args.cpu_cells, cnel, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
# the other BASELINE configs (one GPU): --config p2 | c3 | c4
# --------------------------------------------------------------------------
CONFIGS = {
    "p2": ("ElementTetP2 laplace on MeshTet.init_tensor {n}^3 pts (the metric's P2 case)", 61),
    "c3": ("ElementVector(ElementTetP2) linear_elasticity(lame_parameters(1e3, 0.3)) on "
           "MeshTet.init_tensor {n}^3 pts (BASELINE configs[2])", 70),
    "c4": ("ElementHex2 laplace on MeshHex.init_tensor {n}^3 pts (BASELINE configs[3], "
           "sum-factorised local kernel)", 65),
}


def build_config(name, npts, pkg):
    """(mesh, element, form) of a named config from package `pkg` (skfem_b200 or the
    reference's skfem - same constructors, same names)."""
    x = np.linspace(0, 1, npts)
    if name == "p2":
        from importlib import import_module
        lap = import_module(pkg.__name__ + ".models.poisson").laplace
        return pkg.MeshTet.init_tensor(x, x, x), pkg.ElementTetP2(), lap
    if name == "c3":
        from importlib import import_module
        el = import_module(pkg.__name__ + ".models.elasticity")
        return (pkg.MeshTet.init_tensor(x, x, x), pkg.ElementVector(pkg.ElementTetP2()),
                el.linear_elasticity(*el.lame_parameters(1e3, 0.3)))
    if name == "c4":
        from importlib import import_module
        lap = import_module(pkg.__name__ + ".models.poisson").laplace
        return pkg.MeshHex.init_tensor(x, x, x), pkg.ElementHex2(), lap
    raise ValueError(name)


def run_config(args):
    """Warm re-assembly of a named config through the generic path (element-local kernel ->
    HBM -> deterministic segmented reduction over the cached plan); same JSON contract."""
    import torch
    torch.cuda.set_device(0)
    import skfem_b200 as fem
    from skfem_b200 import _lib
    apply_options(args)
    if args.debug_flags:
        _lib.lib().skb_debug_flags(args.debug_flags)   # kernel variants (A/B runs)
    name = args.config
    npts = args.cells + 1 if args.cells != 100 else CONFIGS[name][1]
    m, elem, form = build_config(name, npts, fem)
    basis = fem.Basis(m, elem)
    nel = m.nelements
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A = form.assemble_device(basis)
    torch.cuda.synchronize()
    cold_ms = 1e3 * (time.perf_counter() - t0)
    nnz = A.nnz
    out = torch.empty(nnz, dtype=torch.float64, device="cuda")

    def step():
        form.assemble_device(basis, out=out)
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    _lib.lib().skb_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    launches = int(_lib.lib().skb_launch_count(0))
    ms_step = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop()
    # size-independent checks: constants / rigid translations in the kernel, symmetry of the sums
    S = torch.sparse_csr_tensor(A.indptr.long(), A.indices.long(), out, size=A.shape)
    v = torch.zeros(basis.N, dtype=torch.float64, device="cuda")
    v[::basis.ncomp] = 1.0
    scale = float(out.abs().max())
    checks = {"max_abs_A_times_constant_over_max_entry": float((S @ v).abs().max()) / scale,
              "nnz": nnz}
    if name == "c4":
        checks["nnz_closed_form"] = (8 * (npts - 1) + 1) ** 3
    checks["ok"] = bool(checks["max_abs_A_times_constant_over_max_entry"] < 1e-10 and
                        checks.get("nnz_closed_form", nnz) == nnz)
    nverts = m.p.shape[1]
    nbs = basis.nbs
    algo = (4 * m.t.shape[0] * nel + 8 * m.p.shape[0] * nverts + 8 * nnz +
            (4 * nbs * nel if basis.element_dofs is not m.t else 0))
    peak, peak_src = peaks()
    achieved = algo / (ms_step * 1e-3) / 1e9
    # end to end before the CPU arm: with the reference (multi-GB numpy temporaries for Hex2)
    # run first, 3 of 4 C4 runs showed 0.2 - 0.4 s stalls per cold call, all inside the plan
    # builder's stream synchronisations (cProfile, profiles/r2_hex_sumfact.md); without the CPU
    # arm, and in tools/c4_e2e_times.py, the same call takes 52 ms
    e2e = None
    if not args.no_e2e and nnz < 2.5e8:
        p_host, t_host = m.p, m.t

        def e2e_step():
            mm = type(m)(p_host, t_host)
            return form.assemble(fem.Basis(mm, elem))
        # the first calls grow torch's pinned-host pool for the result arrays (seconds for the
        # 1.6 GB of C4: tools/c4_e2e_times.py), so four untimed calls like the default config
        for _ in range(4):
            Ah = e2e_step()
        torch.cuda.synchronize()
        k = 3
        if os.environ.get("BENCH_PROFILE_E2E"):       # host-side profile of one call -> stderr
            import cProfile, pstats
            pr = cProfile.Profile()
            pr.enable()
            e2e_step()
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("cumtime").print_stats(25)
        calls = []
        t0 = time.perf_counter()
        for _ in range(k):
            t1 = time.perf_counter()
            Ah = e2e_step()
            calls.append(1e3 * (time.perf_counter() - t1))
        dt = (time.perf_counter() - t0) / k
        e2e = {"value": nel / dt, "unit": "elements/s", "ms_calls": calls,
               "h2d_bytes_per_step": int(m.p.nbytes + m.t.nbytes),
               "d2h_bytes_per_step": int(Ah.data.nbytes + Ah.indices.nbytes + Ah.indptr.nbytes),
               "ms_per_step": 1e3 * dt,
               "what": "cold: Mesh(p,t) + Basis (topology, DOF numbering) + form.assemble -> "
                       "scipy csr_matrix, plan build included"}
    # CPU: the reference itself on a bounded chunk of the same mesh (CellBasis(elements=...))
    cpu = None
    if not args.no_cpu:
        skfem = reference_package()
        if skfem is not None:
            mr, er, fr = build_config(name, npts, skfem)
            chunk = np.arange(min(nel, {"p2": 200000, "c3": 8000, "c4": 1500}[name]))
            t0 = time.perf_counter()
            fr.assemble(skfem.Basis(mr, er, elements=chunk))
            dt = time.perf_counter() - t0
            cpu = {"value": len(chunk) / dt, "unit": "elements/s", "cores": 1,
                   "kind": "reference",
                   "sample": "the unmodified reference (oracle/_ref), single thread, "
                             "Basis(elements=first {} elements) + assemble of the same mesh, "
                             "{:.1f} s".format(len(chunk), dt)}
    line = {
        "metric": "assembly elements/s (FP64, config {})".format(name), "value": nel / (ms_step * 1e-3),
        "unit": "elements/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "nnz_per_s": nnz / (ms_step * 1e-3),
        "config": {"workload": CONFIGS[name][0].format(n=npts) + ", warm re-assembly into CSR",
                   "elements_per_gpu": nel, "dofs_per_gpu": basis.N, "nnz_per_gpu": nnz,
                   "cold_plan_build_ms": cold_ms, "path": "generic",
                   "l2": "no flush: local data and plan exceed the 126 MB L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": algo,
                     "note": "compulsory bytes t + p + scalar element_dofs + CSR data (SURVEY "
                             "8d); the generic path additionally writes and re-reads the "
                             "element-local data"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "checks": checks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "p2", "c3", "c4"],
                    help="c2 (default): BASELINE configs[1], the headline; p2 / c3 / c4: the other "
                         "single-GPU configs through the generic path (N=1 only)")
    ap.add_argument("--cells", type=int, default=100, help="cells per side (per GPU when weak)")
    ap.add_argument("--strong", action="store_true",
                    help="N>1: partition one global init_tensor mesh (cells per side) instead of "
                         "one slab per GPU (BASELINE configs[4] with --cells 250)")
    ap.add_argument("--ref-cells", type=int, default=0, dest="ref_cells",
                    help="reference arm: cells per side of the CPU sample (0 = sized from "
                         "--ref-seconds)")
    ap.add_argument("--ref-seconds", type=float, default=150.0, dest="ref_seconds",
                    help="reference arm: time budget of the K timed steps")
    ap.add_argument("--cpu-cells", type=int, default=60, dest="cpu_cells",
                    help="cpu_baseline leg of the GPU arm: cells per side (60 -> 1.3 M tets)")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e")
    ap.add_argument("--no-traffic", action="store_true", dest="no_traffic",
                    help="skip the ncu child run that measures the step's DRAM traffic")
    ap.add_argument("--no-check", action="store_true", dest="no_check",
                    help="N>1: skip the NCCL parity check against the oracle")
    ap.add_argument("--traffic-child", action="store_true", dest="traffic_child",
                    help=argparse.SUPPRESS)
    ap.add_argument("--static-points", action="store_true", dest="static_points",
                    help="re-assemble the same coordinates every step (round-1 behaviour)")
    ap.add_argument("--debug-flags", type=int, default=0, dest="debug_flags",
                    help="profiling aid (skb_debug_flags): 1 skip P1, 2 skip P2; invalid results")
    ap.add_argument("--no-graph", action="store_true", dest="no_graph",
                    help="launch the warm step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-hex-sumfact", action="store_true", dest="no_hex_sumfact",
                    help="config c4: the Gram-matrix tensor-core kernel instead of sum factorisation")
    ap.add_argument("--no-element-major", action="store_true", dest="no_element_major",
                    help="config c4: local data in the reference layout (Nb, Nb, nel) on warm calls")
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
                    help="v1: do not pin the scratch array in L2 between its two kernels")
    ap.add_argument("--tiling", default="morton", choices=["morton", "kd"],
                    help="v1 fused plan: how elements are cut into tiles")
    ap.add_argument("--fused-version", type=int, default=2, dest="fused_version",
                    help="2: super-tile / pool kernel (default), 1: first-generation kernel")
    ap.add_argument("--tile2", type=int, default=256,
                    help="v2: elements per tile = compute threads per CTA")
    ap.add_argument("--ring2", type=int, default=3, help="v2: record buffers per CTA")
    ap.add_argument("--pool", type=int, default=2048, help="v2: accumulators per super-tile")
    ap.add_argument("--ctas", type=int, default=0, help="v2: CTAs per SM (0 = as many as fit)")
    ap.add_argument("--ept", type=int, default=1,
                    help="v2: elements per compute thread (2 with --tile2 512: 256 threads)")
    ap.add_argument("--super-tiles", type=int, default=None, dest="super_tiles",
                    help="v2: tiles per super-tile (default: as many as the pool allows)")
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--threads", type=int, default=480,
                    help="v1: reduce threads per CTA of the fused kernel")
    args = ap.parse_args()
    if args.traffic_child:
        traffic_child(args)
        return
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference(args)
        return
    if args.config != "c2":
        run_config(args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
