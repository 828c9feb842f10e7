#!/bin/bash
# last 1-GPU session of round 2: the GPU suite, smoke, the default bench line, reductions on views
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
c = d["config"]
print("value %.4g" % d["value"], "ms/iter", c["ms_per_iteration"], "host", c["host_issue_ms_per_iteration"], "host(empty queue)", c.get("host_issue_ms_per_iteration_empty_queue"), "e2e %.4g" % d["e2e"]["value"], "frac", d["roofline"]["frac"])
PY
timeout 60 python benchmarks/red_layouts.py > gpurun_out/r02_red_layouts.jsonl 2> gpurun_out/r02_red_layouts.err; echo "layouts rc=$?"; cat gpurun_out/r02_red_layouts.jsonl; tail -3 gpurun_out/r02_red_layouts.err
