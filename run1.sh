set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_binary_red.py tests/test_parity_reductions.py -x -q -m gpu > gpurun_out/t_new.log 2>&1; echo "new rc=$?" >> gpurun_out/t_new.log
tail -5 gpurun_out/t_new.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
python benchmarks/sweep.py --c3 --reps 10 > gpurun_out/c3_tma.jsonl 2> gpurun_out/c3_tma.err
CNB_AXIS_ROW_TMA=0 python benchmarks/sweep.py --c3 --reps 10 > gpurun_out/c3_ldg.jsonl 2> gpurun_out/c3_ldg.err
python benchmarks/sweep.py --c5 --reps 10 > gpurun_out/c5.jsonl 2> gpurun_out/c5.err
python bench.py > gpurun_out/bench_bs.json 2> gpurun_out/bench_bs.err
python bench.py --workload stencil > gpurun_out/bench_stencil.json 2> gpurun_out/bench_stencil.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
grep -h "axis1" gpurun_out/c3_tma.jsonl gpurun_out/c3_ldg.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['case'], round(r['ms'],3), round(r['frac_of_hbm_peak_per_gpu'],3))
"
cat gpurun_out/bench_bs.json
