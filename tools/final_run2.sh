#!/bin/bash
# evidence refresh after the Hex2 sum factorisation / element-major warm path
bash tools/final_run.sh
bash tools/ncu_tour.sh 41 420
HEXSF_ARGS="" bash tools/hex_sf_ncu.sh > gpurun_out/hexsf_ncu_final.txt 2>&1
tail -8 gpurun_out/hexsf_ncu_final.txt | cut -c1-160
timeout 300 ncu --set full --clock-control none --kernel-id "::regex:csr_reduce_em_kernel:1" -o gpurun_out/reduce_em_c3 -f \
  python bench.py --config c3 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/reduce_em_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
