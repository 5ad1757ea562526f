#!/bin/bash
# GPU check of the sum-factorised Hex2 kernel: parity tests, then C4 with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plan_rows.py -q -m gpu -k "hex" > gpurun_out/hexsf_tests.txt 2>&1
tail -5 gpurun_out/hexsf_tests.txt
timeout 300 python bench.py --config c4 --no-cpu --no-e2e --steps 5 > gpurun_out/hexsf_c4_on.json 2> gpurun_out/hexsf_c4_on.err
timeout 300 python bench.py --config c4 --no-cpu --no-e2e --steps 5 --no-hex-sumfact > gpurun_out/hexsf_c4_off.json 2> gpurun_out/hexsf_c4_off.err
timeout 300 python bench.py --config c4 --no-cpu --no-e2e --steps 5 --debug-flags 16 > gpurun_out/hexsf_c4_v16.json 2> gpurun_out/hexsf_c4_v16.err
timeout 300 python bench.py --config c4 --no-cpu --no-e2e --steps 5 --no-element-major > gpurun_out/hexsf_c4_noem.json 2> gpurun_out/hexsf_c4_noem.err
python - <<'P'
import json
for n in ("on","off","v16","noem"):
    try:
        d=json.loads(open(f"gpurun_out/hexsf_c4_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["checks"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/hexsf_c4_{n}.err").read()[-1500:])
P
