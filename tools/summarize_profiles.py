"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked
summaries under profiles/ (round-1 naming).  Usage:
    python tools/summarize_profiles.py r1
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. launch list -> per-kernel average duration and share, warm step and one-time work apart
src = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")))
            for r in csv.DictReader(lines)]
    first = next((i for i, (n, _) in enumerate(rows) if "p1tet_laplace_fused" in n), len(rows))

    def table(f, part):
        agg = collections.OrderedDict()
        for n, v in part:
            agg.setdefault(n, []).append(v)
        tot = sum(sum(v) for v in agg.values()) or 1.0
        f.write("| kernel | launches | avg us | share of listed time |\n|---|---|---|---|\n")
        for k, v in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k[:110], len(v), sum(v) / len(v) / 1e3,
                                                         100 * sum(v) / tot))
        return tot

    with open(os.path.join(P, "%s_launches.md" % tag), "w") as f:
        f.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n"
                "Command: `python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-graph` (1 GPU, "
                "%d launches in total). Durations are cold-cache and serialised under ncu; compare "
                "shares.\n\n## Warm step (the timed region: every launch from the first fused "
                "launch on)\n\n" % (tag, len(rows)))
        table(f, rows[first:])
        lib = [(n, v) for n, v in rows[:first] if "skb::" in n]
        other = [("torch / CUB kernels of the cold plan build and the fused-plan build "
                  "(sort, unique, searchsorted, index ops)", v) for n, v in rows[:first]
                 if "skb::" not in n]
        f.write("\n## One-time work before it (cold assembly + plan builds)\n\n")
        table(f, lib + other)
    print("wrote launches")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_sample_buffer_full"]
dram_total = 0.0
for name in ("fused", "combine"):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    with open(os.path.join(P, "%s_ncu_%s.md" % (tag, name)), "w") as f:
        f.write("# ncu --set full --clock-control none: `%s`\n\n" % d.get("Kernel Name", ("?",))[0])
        f.write("Command: `python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph` "
                "(BASELINE configs[1], 6.0 M P1 tets, 1 GPU). One launch.\n\n| metric | value | unit |\n|---|---|---|\n")
        for w in WANT:
            if w in d:
                f.write("| %s | %s | %s |\n" % (w, d[w][0], d[w][1]))
        try:
            rd = float(d["dram__bytes_read.sum"][0]); wr = float(d["dram__bytes_write.sum"][0])
            f.write("\nDRAM traffic of this launch: %.1f MB (read %.1f + write %.1f).\n" % (rd + wr, rd, wr))
            mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            dram_total += rd * mult.get(d["dram__bytes_read.sum"][1], 1e6) \
                + wr * mult.get(d["dram__bytes_write.sum"][1], 1e6)
        except Exception:
            pass
    print("wrote", name)
if dram_total:
    with open(os.path.join(P, "traffic.json"), "w") as f:
        json.dump({"cells": 100, "path": "fused", "dram_bytes_per_step": int(dram_total),
                   "source": "profiles/%s_ncu_fused.md + %s_ncu_combine.md "
                             "(dram__bytes_read.sum + dram__bytes_write.sum, one launch each)"
                             % (tag, tag)}, f)
    print("wrote traffic.json", dram_total)
