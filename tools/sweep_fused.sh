#!/bin/bash
# sweep fused-kernel configurations (tile, reduce threads, record ring) on one
# B200 (run under gpurun); prints warm step time, roofline fraction, plan stats.
# usage: tools/sweep_fused.sh ["T R N" ...]   (default: a standard list)
if [ $# -gt 0 ]; then cfgs=("$@"); else
cfgs=("512 256 4" "512 256 5" "512 384 4" "512 480 4" "512 128 4" "256 256 4" "768 224 4"); fi
for cfg in "${cfgs[@]}"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --tile $1 --threads $2 --ring $3 ${4:+--debug-flags $4} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
fp=d['config']['fused_plan']
print('tile',fp['tile'],'red',fp['threads'],'ring',fp['ring'],'dbg','${4:-0}','ms',round(d['ms_per_step'],4),'el/s %.3e'%d['value'],'frac',round(d['roofline']['frac'],4),'B/el',round(fp['per_element'],1),'dup',round(fp['tile_slots_per_csr_slot'],2),'vcap',fp['vcap'],'smem',fp['smem'])
"
done
