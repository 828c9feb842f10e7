#!/bin/bash
# halo exchange overlapped with the interior tiles (fusion.Overlap): correctness at N ranks, then the
# stencil bench with the overlap on and off.  usage: r02_overlap.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -5
for ov in 1 0; do
  CUNUMERIC_B200_HALO_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 5 --no-extras --no-e2e \
    --no-cpu-baseline > gpurun_out/r02_overlap_${N}gpu_ov$ov.json 2> gpurun_out/r02_overlap_${N}gpu_ov$ov.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_overlap_${N}gpu_ov$ov.json").read().strip().splitlines()[-1])
    print("overlap=$ov N=$N ms/iter", d["ms_per_step"], "value", d["value"], "host_issue", d.get("host_issue_ms_per_iteration"), d.get("fusion_stats"))
except Exception as e:
    print("overlap=$ov failed", e); print(open("gpurun_out/r02_overlap_${N}gpu_ov$ov.err").read()[-3000:])
PY
done
