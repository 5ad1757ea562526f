#!/bin/bash
# sweep fused-kernel tile/thread configurations on one B200 (run under gpurun)
for cfg in "1024 256" "1024 512" "1024 1024" "2048 512" "2048 1024" "512 256" "512 512"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 3 --no-cpu --tile $1 --threads $2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
fp=d['config']['fused_plan']
print('tile',fp['tile'],'threads',fp['threads'],'ms',round(d['ms_per_step'],4),'el/s %.3e'%d['value'],'frac',round(d['roofline']['frac'],4),'B/el',round(fp['per_element'],1),'dup',round(fp['tile_slots_per_csr_slot'],2),'vcap',fp['vcap'],'sellpad',round(fp['sell_padding'],3))
"
done
