#!/bin/bash
# A/B of several builds of the library on ONE box (box-to-box spread makes numbers from
# different gpurun calls incomparable): runs the same bench command in every directory given,
# round-robin, twice.  usage: tools/ab_run.sh "<bench flags>" dir1 dir2 ...
flags=$1; shift
for rep in 1 2; do
  for d in "$@"; do
    ms=$(cd $d && timeout 200 python bench.py --no-e2e --no-cpu --no-traffic --steps 20 --warmup 3 $flags 2>/dev/null \
      | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('%.4f ms  clocks %s  ok %s' % (d['ms_per_step'], d.get('clocks', {}).get('sm_mhz'), (d.get('checks') or {}).get('ok')))")
    echo "$d [$flags] rep $rep: $ms"
  done
done
