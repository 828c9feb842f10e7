mkdir -p gpurun_out
nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -15
N=$(nvidia-smi -L | wc -l)
for n in 1 $N; do
timeout 900 python bench.py --gpus $n --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err; echo "bench $n rc=$?"
tail -3 gpurun_out/bench_g$n.err
python -c "
import sys, json
d = json.loads(open('gpurun_out/bench_g$n.json').read().strip().splitlines()[-1]); print('N=$n', d['config']['ms_per_iteration'], d['roofline']['frac'], d['value'], d['e2e'] and d['e2e']['value'], d['roofline']['per_kernel'])"
done
