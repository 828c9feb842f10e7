mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
nvidia-smi topo -m | head -12
for n in 8 4; do
timeout 600 python bench.py --gpus $n --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err; echo "bench $n rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_g$n.err | tail -4
python -c "
import sys, json
d = json.loads(open('gpurun_out/bench_g$n.json').read().strip().splitlines()[-1]); print('N=$n', d['config']['ms_per_iteration'], d['roofline']['frac'], d['value'], d['e2e'] and d['e2e']['value'], d['roofline']['per_kernel'], d['host_binding'])"
done
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -4
