#!/bin/bash
# one session on an 8-GPU box: distributed tests, halo-overlap on/off at 8 ranks, the record runs
mkdir -p gpurun_out
python -m pytest tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -4
best=1; best_ms=1e9
for ov in 1 0 1 0; do
  CUNUMERIC_B200_HALO_OVERLAP=$ov timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-e2e \
    --no-cpu-baseline > gpurun_out/r02_ov8_$ov.json 2> gpurun_out/r02_ov8_$ov.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_ov8_$ov.json").read().strip().splitlines()[-1])
print("overlap=$ov N=8 ms/iter", d["config"]["ms_per_iteration"], "value %.4g" % d["value"], "host", d["config"].get("host_issue_ms_per_iteration"), "launches", d["gpu_launches"])
PY
done
for n in 8 4; do
  timeout 600 python bench.py --gpus $n --steps 10 --warmup 5 > gpurun_out/r02_bench_default_${n}gpu.json 2> gpurun_out/r02_bench_default_${n}gpu.err; echo "bench $n rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_default_${n}gpu.json").read().strip().splitlines()[-1])
print("N=$n", "%.4g" % d["value"], d["config"]["ms_per_iteration"], d["config"]["host_issue_ms_per_iteration"], "e2e %.4g" % d["e2e"]["value"], d["config"].get("fusion_stats"))
for c in d["sweeps"]["cases"]: print(c["case"], c["ms"], c["gbs_per_gpu"], c["frac"])
PY
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  benchmarks/pcie_ranks.py 2>/dev/null | grep '^{' | tee gpurun_out/r02_pcie_8ranks.json
