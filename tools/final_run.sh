#!/bin/bash
# End-of-round evidence on one B200: GPU tests, smoke, the default bench line, the other
# configs, launch lists.  Everything under its own timeout.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2_gpu_tests_final.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 300 gpurun_out/r2_bench_final.json
for c in p2 c3 c4; do
  timeout 280 python bench.py --config $c --steps 5 --warmup 3 2>gpurun_out/r2_cfg_$c.err > gpurun_out/r2_cfg_$c.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_cfg_$c.json").read().strip().splitlines()[-1])
print("$c", d["ms_per_step"], d["value"], d["roofline"]["frac"], (d.get("e2e") or {}).get("ms_per_step"), d["checks"]["ok"], (d.get("cpu_baseline") or {}).get("value"))
PY
done
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e \
  --no-traffic --no-graph > gpurun_out/r2_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches.csv | grep -v "at::\|native::\|at_cuda" | tail -30
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_cold_launches_rows.csv python tools/cold_once.py rows > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_cold_launches_rows.csv | tail -22
