mkdir -p gpurun_out/g
python bench.py > gpurun_out/g/bench_bs.json 2> gpurun_out/g/bench_bs.err; tail -2 gpurun_out/g/bench_bs.err
python bench.py --workload stencil > gpurun_out/g/bench_stencil.json 2> gpurun_out/g/bench_stencil.err; tail -2 gpurun_out/g/bench_stencil.err
python bench.py --workload stencil --fusion off --steps 5 > gpurun_out/g/bench_stencil_obo.json 2> /dev/null
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 40 -c 40 --csv --log-file gpurun_out/g/bs_fused_launches.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g/l.log 2>&1
$NCU --set full --import-source on -k regex:fused_ -s 4 -c 1 -o gpurun_out/g/bs_fused python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g/n1.log 2>&1
ncu -i gpurun_out/g/bs_fused.ncu-rep --page raw --csv > gpurun_out/g/bs_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/g/bs_fused.ncu-rep --page details --csv > gpurun_out/g/bs_fused_details.csv 2>/dev/null
$NCU --set full -k regex:fused_ -s 40 -c 1 -o gpurun_out/g/st_fused python bench.py --workload stencil --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/g/n2.log 2>&1
ncu -i gpurun_out/g/st_fused.ncu-rep --page raw --csv > gpurun_out/g/st_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/g/st_fused.ncu-rep --page details --csv > gpurun_out/g/st_fused_details.csv 2>/dev/null
rm -f gpurun_out/g/*.ncu-rep
python - <<'PY'
import json
r = json.load(open('gpurun_out/g/bench_bs.json'))
print('BS', r['value'], r['ms_per_step'], r['gpu_launches'], r['roofline']['kernel'], r['roofline']['frac'], r['roofline']['avg_launch_ms'], 'obo', r['op_by_op']['value'], 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'], r['clocks'])
for f in ('bench_stencil', 'bench_stencil_obo'):
    r = json.load(open(f'gpurun_out/g/{f}.json'))
    print('ST', r['value'], r['ms_per_step'], r['gpu_launches'], r['roofline'].get('per_kernel'), r['whole_iteration']['frac_of_hbm_peak'])
PY
