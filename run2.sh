mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_reductions.py -x -q -m gpu > gpurun_out/t_red.log 2>&1; echo "rc=$?" >> gpurun_out/t_red.log
tail -3 gpurun_out/t_red.log
python benchmarks/sweep.py --c3 --reps 10 > gpurun_out/c3_tma3.jsonl 2> gpurun_out/c3_tma3.err
for c in 2 4 6; do CNB_EW_CTAS_PER_SM=$c python benchmarks/sweep.py --fill --reps 10 > gpurun_out/fill_c$c.jsonl 2> gpurun_out/fill.err; done
for f in gpurun_out/c3_tma3.jsonl gpurun_out/fill_c*.jsonl; do echo $f; cat $f | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['case'], round(r['ms'],3), round(r['algorithmic_gbs_per_gpu']), round(r['frac_of_hbm_peak_per_gpu'],3))
"; done
