#!/bin/bash
# reductions of misaligned pitched views (head | body | tail): the GPU suite, then the layout table
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 60 python benchmarks/red_layouts.py > gpurun_out/r02_red_layouts_peel.jsonl 2> gpurun_out/r02_red_layouts_peel.err; echo "layouts rc=$?"; cat gpurun_out/r02_red_layouts_peel.jsonl; tail -3 gpurun_out/r02_red_layouts_peel.err
