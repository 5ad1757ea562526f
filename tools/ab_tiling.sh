#!/bin/bash
# A/B of the fused plan's tiling on one B200 (first GPU call of round 2):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/ab_tiling.sh'
# 1. parity of the opt-in k-d plan through the C ABI (tests/test_gpu_parity.py)
# 2. warm-step time of BASELINE configs[1] for Morton / k-d / k-d with a 5-deep record ring,
#    interleaved so that box-to-box and thermal drift show up as a spread between equal runs
# Every line lands in gpurun_out/ab_tiling.jsonl; nothing here runs under a profiler.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/ab_tiling.jsonl
: > "$out"
SKB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q \
    -k "kd_tiling or fused_p1_path or determinism" > gpurun_out/ab_tiling_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/ab_tiling_tests.log
tail -3 gpurun_out/ab_tiling_tests.log
for variant in "--tiling morton" "--tiling kd" "--tiling kd --ring 5" \
               "--tiling morton" "--tiling kd" "--tiling kd --ring 5" \
               "--tiling kd --arith fast"; do
    echo "== $variant"
    # shellcheck disable=SC2086
    timeout 600 python bench.py --steps 300 --warmup 30 --no-cpu --no-e2e $variant \
        2> gpurun_out/ab_tiling.err | tail -1 | tee -a "$out" \
        | python -c "import sys, json; d = json.loads(sys.stdin.read()); \
print(d['ms_per_step'], d['config'].get('fused_plan', {}).get('tiling'), \
d['config'].get('fused_plan', {}).get('ring'), d['roofline']['frac'])"
done
