mkdir -p gpurun_out
run() { echo "== $1"; env $1 python bench.py --gpus 8 --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   ms/iter', round(d['config']['ms_per_iteration'], 4), 'host', round(d['config']['host_issue_ms_per_iteration'], 4), 'value', d['value'])"; }
run "X=1"
run "NCCL_PROTO=LL"
run "NCCL_PROTO=LL128"
run "NCCL_MAX_NCHANNELS=2"
