mkdir -p gpurun_out/g8
timeout 300 python -m pytest tests/test_distributed_gpu.py -q -m gpu -k "4 or 8" > gpurun_out/g8/dist.log 2>&1; echo "rc=$?" >> gpurun_out/g8/dist.log
tail -4 gpurun_out/g8/dist.log
timeout 200 python bench.py --gpus 8 > gpurun_out/g8/bench_bs_8gpu.json 2> gpurun_out/g8/bench_bs_8gpu.err
timeout 200 python bench.py --gpus 8 --workload stencil --steps 10 > gpurun_out/g8/bench_stencil_8gpu.json 2> gpurun_out/g8/bench_stencil_8gpu.err
timeout 200 python bench.py --gpus 4 --workload stencil --steps 10 > gpurun_out/g8/bench_stencil_4gpu.json 2> gpurun_out/g8/bench_stencil_4gpu.err
timeout 200 python bench.py --gpus 8 --workload stencil --steps 5 --fusion off > gpurun_out/g8/bench_stencil_8gpu_obo.json 2> /dev/null
python - <<'PY'
import json
for f in ('bench_bs_8gpu', 'bench_stencil_8gpu', 'bench_stencil_4gpu', 'bench_stencil_8gpu_obo'):
    try:
        txt = [l for l in open(f'gpurun_out/g8/{f}.json') if l.startswith('{')][-1]
        r = json.loads(txt)
        print(f, r['value'], r['ms_per_step'], r['gpu_launches'], (r.get('op_by_op') or {}).get('value'), (r.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
