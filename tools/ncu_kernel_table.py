"""One row per kernel from an `ncu --set full` report: the metrics the roofline discussion in
DESIGN.md cites.  usage: ncu_kernel_table.py report.ncu-rep out.md "<title>" """
import csv, subprocess, sys, collections
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {k: i for i, k in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU %")]

def val(r, name):
    if name not in ix:
        return None
    try:
        v = float(r[ix[name]].replace(",", ""))
    except ValueError:
        return None
    u = units[ix[name]]
    if name.endswith("duration.sum"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
    if "bytes" in name:
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
    return v

agg = collections.OrderedDict()
for r in body:
    name = r[ix["Kernel Name"]].split("(")[0]
    name = name.replace("void ", "").replace("skb::", "")
    agg.setdefault(name, []).append(r)
with open(out, "w") as f:
    f.write("# %s\n\n`ncu --set full --clock-control none` (each kernel replayed in isolation, cold "
            "caches): first launch of every kernel of the library in `tools/kernel_tour.py`; where a "
            "kernel is launched several times the largest launch is shown.\n\n" % title)
    f.write("| kernel | launches | " + " | ".join(c for _, c in COLS) + " |\n")
    f.write("|---|---|" + "---|" * len(COLS) + "\n")
    for name, rs in agg.items():
        r = max(rs, key=lambda r: val(r, "gpu__time_duration.sum") or 0)
        cells = []
        for m, _ in COLS:
            v = val(r, m)
            cells.append("-" if v is None else ("%.0f" % v if v >= 100 or float(v).is_integer() else "%.1f" % v))
        f.write("| `%s` | %d | %s |\n" % (name[:70], len(rs), " | ".join(cells)))
print("wrote", out, len(agg), "kernels")
