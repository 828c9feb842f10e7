mkdir -p gpurun_out/g2
timeout 600 python -m pytest tests/test_distributed_gpu.py -x -q -m gpu > gpurun_out/g2/dist.log 2>&1; echo "rc=$?" >> gpurun_out/g2/dist.log
tail -15 gpurun_out/g2/dist.log
python bench.py --gpus 2 > gpurun_out/g2/bench_bs_2gpu.json 2> gpurun_out/g2/bench_bs_2gpu.err
python bench.py --gpus 2 --workload stencil > gpurun_out/g2/bench_stencil_2gpu.json 2> gpurun_out/g2/bench_stencil_2gpu.err
python - <<'PY'
import json
for f in ('bench_bs_2gpu', 'bench_stencil_2gpu'):
    try:
        r = json.load(open(f'gpurun_out/g2/{f}.json'))
        print(f, r['value'], r['ms_per_step'], r['gpu_launches'], (r.get('op_by_op') or {}).get('value'), (r.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/g2/{f}.err').read()[-1500:])
PY
