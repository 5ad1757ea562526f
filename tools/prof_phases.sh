# per-phase cycle counters of the fused kernel (debug bit 2: CTA 0 prints, per warp)
for c in 1 3; do
  echo "== ctas $c"
  python bench.py --no-e2e --no-cpu --no-traffic --no-check --steps 1 --warmup 3 --ctas $c --debug-flags 4 "$@" 2>/dev/null | grep -E "fused2 cta0" | tail -9 | cut -c1-300
done
