set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion.py -m gpu -x -q 2>&1 | tail -3
for cs in 1 0; do
CUNUMERIC_B200_TMA_CSHIFT=$cs timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2> gpurun_out/bench_stencil.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CSHIFT=$cs', d['config']['ms_per_iteration'], d['roofline']['frac'], d['value'])"
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -k regex:fused_ -s 20 -c 2 --csv python bench.py --steps 1 --warmup 3 --stencil-iters 10 --no-extras --no-e2e --no-cpu-baseline 2>&1 | grep -v "^==" | cut -d, -f5,13- | tail -9
