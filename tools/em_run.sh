#!/bin/bash
# element-major warm path: parity tests, then p2 / c3 / c4 with and without it on one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "element_major or hex" > gpurun_out/em_tests.txt 2>&1
tail -5 gpurun_out/em_tests.txt
for c in p2 c3 c4; do
  timeout 500 python bench.py --config $c --no-cpu --no-e2e --steps 5 > gpurun_out/em_${c}_on.json 2> gpurun_out/em_${c}_on.err
  timeout 500 python bench.py --config $c --no-cpu --no-e2e --steps 5 --no-element-major > gpurun_out/em_${c}_off.json 2> gpurun_out/em_${c}_off.err
done
python - <<'P'
import json
for c in ("p2","c3","c4"):
  for n in ("on","off"):
    try:
        d=json.loads(open(f"gpurun_out/em_{c}_{n}.json").read().strip().splitlines()[-1])
        print(c, n, d["ms_per_step"], d["checks"]["ok"], d["gpu_launches"])
    except Exception as e:
        print(c, n, "failed", e); print(open(f"gpurun_out/em_{c}_{n}.err").read()[-1500:])
P
