mkdir -p gpurun_out/v
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/v/all.log 2>&1; echo "rc=$?" >> gpurun_out/v/all.log
tail -4 gpurun_out/v/all.log
python bench.py > gpurun_out/v/bench_bs.json 2> gpurun_out/v/bench_bs.err
python bench.py --workload stencil --steps 20 > gpurun_out/v/bench_stencil.json 2> gpurun_out/v/bench_stencil.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
r = json.loads([l for l in open('gpurun_out/v/bench_bs.json') if l.startswith('{')][-1])
print('BS', r['value'], r['ms_per_step'], r['gpu_launches'], r['roofline']['frac'], r['roofline']['avg_launch_ms'], 'obo', r['op_by_op']['value'], r['op_by_op']['roofline']['frac'], 'e2e', r['e2e']['value'], r['clocks'], 'cpu', r['cpu_baseline']['value'])
r = json.loads([l for l in open('gpurun_out/v/bench_stencil.json') if l.startswith('{')][-1])
print('ST', r['value'], r['ms_per_step'], r['gpu_launches'])
PY
