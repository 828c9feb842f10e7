mkdir -p gpurun_out
for p in 0 128; do PROMO=$p ./benchmarks/micro/stencil_tma 40000 20 | grep -E "TMA|promo"; done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
for c in 3 7; do
PROMO=0 HINT=1 CONFIG=$c ncu --metrics $M --clock-control none -k regex:stencil_tma -s 5 -c 1 --csv ./benchmarks/micro/stencil_tma 40000 3 2>&1 | grep -v "^==" | tail -4 | cut -d, -f5,13-
done
