mkdir -p gpurun_out
for n in 8; do
timeout 900 python bench.py --gpus $n --steps 10 --warmup 5 > gpurun_out/r02_bench_default_${n}gpu.json 2> gpurun_out/r02_bench_default_${n}gpu.err; echo "bench $n rc=$?"
grep -v "^\*\|OMP_NUM\|^$\|NCCL communicator" gpurun_out/r02_bench_default_${n}gpu.err | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_default_${n}gpu.json").read().strip().splitlines()[-1])
print("N=$n", d["value"], d["config"]["ms_per_iteration"], d["config"]["host_issue_ms_per_iteration"], "e2e", d["e2e"]["value"], d["host_binding"])
bs=d["black_scholes"]; print("BS", bs["value"], bs["ms_per_step"], bs["roofline"]["avg_launch_ms"], "e2e", bs["e2e"]["value"])
for c in d["sweeps"]["cases"]: print(c["case"], c["ms"], c["gbs_per_gpu"], c["frac"])
PY
done
