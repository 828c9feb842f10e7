#!/bin/bash
# full reduction of pitched views through the row kernels: reduction suites, layout table, and the
# same view tests with the new route switched off (the fallback)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_parity_reductions.py tests/test_api.py "tests/test_parity_config_scale.py::test_c3_reductions_32768_squared" -m gpu -x -q 2>&1 | tail -4
timeout 40 python benchmarks/red_layouts.py > gpurun_out/r02_red_layouts_peel.jsonl 2> gpurun_out/r02_red_layouts_peel.err; echo "layouts rc=$?"; grep "pitched\|every other row\|contiguous\|match" gpurun_out/r02_red_layouts_peel.jsonl; tail -3 gpurun_out/r02_red_layouts_peel.err
CNB_RED_PEEL_SCALAR=0 timeout 40 python -m pytest tests/test_parity_reductions.py -m gpu -x -q -k "misaligned or views" 2>&1 | tail -3
