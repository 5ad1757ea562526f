#!/bin/bash
# A/B sweep of the second-generation fused kernel's launch shape on one B200 (run under gpurun).
# Each line: tile ring pool ctas [extra bench flags]
mkdir -p gpurun_out
out=gpurun_out/sweep_fused2.txt
: > $out
run() {
  echo "== $*" >> $out
  python bench.py --no-e2e --no-cpu --steps 20 --warmup 3 "$@" 2>>gpurun_out/sweep_fused2.err \
    | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    fp = d['config'].get('fused_plan') or {}
    print('ms_per_step %.4f frac %.4f smem %s S %s pool %s shared %s recB/el %.1f total_B/el %.1f' % (
        d['ms_per_step'], d['roofline']['frac'], fp.get('smem'), fp.get('super_tile_tiles'),
        fp.get('pool_cap'), fp.get('shared_slots'), fp.get('records', 0) / d['config']['elements_per_gpu'],
        fp.get('per_element', 0)))
" >> $out
}
# configurations: the file given as $1, else the default list below
cfg=${1:-}
if [ -z "$cfg" ]; then
  cfg=$(mktemp)
  cat > $cfg <<'CFG'
--tile2 256 --ring2 3 --pool 2048
--tile2 256 --ring2 4 --pool 2048
--tile2 256 --ring2 3 --pool 4096
--tile2 512 --ring2 3 --pool 4096
--tile2 512 --ring2 3 --pool 2048
--tile2 128 --ring2 3 --pool 2048
--tile2 256 --ring2 3 --pool 2048 --debug-flags 1
--tile2 256 --ring2 3 --pool 2048 --debug-flags 2
--tile2 256 --ring2 3 --pool 2048 --debug-flags 3
--tile2 256 --ring2 3 --pool 2048 --debug-flags 67
--tile2 256 --ring2 3 --pool 2048 --arith fast
CFG
fi
[ -r "$cfg" ] || { echo "no such configuration file: $cfg" >&2; exit 2; }
while read -r line; do
  [ -z "$line" ] && continue
  run $line
done < "$cfg"
cat $out
