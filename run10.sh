mkdir -p gpurun_out/h
for c in 8 16 32; do
  python bench.py --steps 10 --no-cpu-baseline --e2e-chunks $c 2>/dev/null | python -c "
import sys, json
r = json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('chunks $c', r['value'], r['ms_per_step'], r['e2e']['value'], r['e2e']['ms_per_step'], r['clocks']['samples'])"
done
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:fused_ -c 13 --csv --log-file gpurun_out/h/bs_fused_launches.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/h/bs_fused_launches.csv
