set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for c in 3 4 5 6 8; do for u in "" 1 4; do
  echo "== ctas=$c U=$u"
  CNB_FUSED_CTAS_PER_SM=$c CNB_FUSED_U=$u python bench.py --workload black_scholes --steps 30 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('RESULT', d['ms_per_step'], d['value'], d['roofline']['frac'])
"
done; done
echo "== stencil baseline"
python bench.py --workload stencil --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-600
