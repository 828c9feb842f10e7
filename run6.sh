mkdir -p gpurun_out/f
timeout 900 python -m pytest tests/test_fusion.py -q -m gpu > gpurun_out/f/fusion.log 2>&1; echo "rc=$?" >> gpurun_out/f/fusion.log
tail -30 gpurun_out/f/fusion.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f/all.log 2>&1; echo "rc=$?" >> gpurun_out/f/all.log
tail -5 gpurun_out/f/all.log
python bench.py > gpurun_out/f/bench_bs.json 2> gpurun_out/f/bench_bs.err; tail -3 gpurun_out/f/bench_bs.err
python bench.py --workload stencil > gpurun_out/f/bench_stencil.json 2> gpurun_out/f/bench_stencil.err; tail -3 gpurun_out/f/bench_stencil.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f/smoke.log 2>&1; tail -2 gpurun_out/f/smoke.log
python - <<'PY'
import json
r = json.load(open('gpurun_out/f/bench_bs.json'))
print('BS', r['value'], r['ms_per_step'], r['gpu_launches'], r['roofline']['kernel'], r['roofline']['frac'], r['roofline']['achieved'], 'obo', r['op_by_op']['value'], 'e2e', r['e2e'])
r = json.load(open('gpurun_out/f/bench_stencil.json'))
print('ST', r['value'], r['ms_per_step'], r['gpu_launches'], r['roofline'].get('per_kernel'), r['whole_iteration'])
PY
cat cunumeric_b200/_fused_cache/*.err 2>/dev/null | head -50
