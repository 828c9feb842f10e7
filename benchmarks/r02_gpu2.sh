mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion.py tests/test_api.py -m gpu -x -q 2>&1 | tail -3
for tr in 16 32 8; do
CNB_TMA_TR=$tr timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2> gpurun_out/bench_stencil.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TR=$tr', d['config']['ms_per_iteration'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['value'], d['config']['fusion_stats'])"
tail -3 gpurun_out/bench_stencil.err
done
