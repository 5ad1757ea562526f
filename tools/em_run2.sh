#!/bin/bash
# reduce kernels with side-by-side fetches: full GPU suite, then c3 / p2 / c4 / c2-generic timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/em2_tests.txt 2>&1
tail -4 gpurun_out/em2_tests.txt
for c in p2 c3 c4; do
  timeout 500 python bench.py --config $c --no-cpu --no-e2e --steps 5 > gpurun_out/em2_${c}.json 2> gpurun_out/em2_${c}.err
done
python - <<'P'
import json
for c in ("p2","c3","c4"):
    try:
        d=json.loads(open(f"gpurun_out/em2_{c}.json").read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["checks"]["ok"], d["gpu_launches"])
    except Exception as e:
        print(c, "failed", e); print(open(f"gpurun_out/em2_{c}.err").read()[-1500:])
P
