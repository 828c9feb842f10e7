# round-2 profile session: launch lists + ncu --set full of the two headline fused kernels
mkdir -p gpurun_out/prof
B="python bench.py --steps 2 --warmup 3 --stencil-iters 10 --no-extras --no-e2e --no-cpu-baseline"
# 1. launch list of the stencil bench command (cold-cache, serialised per-launch times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/r02_stencil_launches_ncu.csv $B > /dev/null 2>&1
# 2. full capture of the TMA stencil kernel (skip warm-up launches)
ncu --set full --import-source on --clock-control none -k regex:fused_.*_tma -s 30 -c 1 -o gpurun_out/prof/r02_fused_stencil_tma $B > /dev/null 2>&1
# 3. Black-Scholes fused kernel
BB="python bench.py --workload black_scholes --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/r02_black_scholes_launches_ncu.csv $BB > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:fused_.*_vec -s 3 -c 1 -o gpurun_out/prof/r02_fused_black_scholes $BB > /dev/null 2>&1
ls -la gpurun_out/prof
for r in r02_fused_stencil_tma r02_fused_black_scholes; do
  ncu -i gpurun_out/prof/$r.ncu-rep --page raw --csv > gpurun_out/prof/${r}_ncu_full_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof/$r.ncu-rep --page source --csv > gpurun_out/prof/${r}_ncu_source.csv 2>/dev/null
done
ls -la gpurun_out/prof
