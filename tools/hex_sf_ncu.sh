#!/bin/bash
# ncu capture of the sum-factorised Hex2 kernel on C4 (first launch) + durations of the warm step's kernels
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on --kernel-id "::regex:local_hex_sf_kernel:1" \
  -o gpurun_out/hexsf_c4 -f python bench.py --config c4 --no-cpu --no-e2e --steps 1 --warmup 1 $HEXSF_ARGS > gpurun_out/hexsf_ncu.log 2>&1
tail -2 gpurun_out/hexsf_ncu.log | cut -c1-300
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:local_hex|csr_reduce|rows_|scan_" --csv --log-file gpurun_out/hexsf_launches.csv \
  python bench.py --config c4 --no-cpu --no-e2e --steps 1 --warmup 1 $HEXSF_ARGS > gpurun_out/hexsf_launches.log 2>&1
cut -d, -f5,12- gpurun_out/hexsf_launches.csv | cut -c1-200 | tail -25
ls -la gpurun_out/hexsf_c4.ncu-rep
