# Black-Scholes fused kernel: chunks per thread (U) x resident CTAs per SM (launch bound)
for u in 1 2; do for mb in 4 5 6; do
  echo "== U=$u minblocks=$mb"
  CNB_FUSED_U=$u CNB_FUSED_VEC_MINBLOCKS=$mb CNB_FUSED_CTAS_PER_SM=8 python bench.py --workload black_scholes --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('RESULT', round(d['ms_per_step'], 4), round(d['roofline']['avg_launch_ms'], 4), d['roofline']['frac'])
"
done; done
