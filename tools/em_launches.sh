#!/bin/bash
# durations of the warm step's kernels (ncu, serialised) for c3 and p2
mkdir -p gpurun_out
for c in c3 p2; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:local_|csr_reduce" --csv --log-file gpurun_out/em_launches_$c.csv \
  python bench.py --config $c --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/em_launches_$c.log 2>&1
python - $c <<'P'
import csv,sys
c=sys.argv[1]
rows=[r for r in csv.reader(open(f"gpurun_out/em_launches_{c}.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
for k,v in d.items(): print(c,k,v)
P
done
