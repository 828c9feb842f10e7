mkdir -p gpurun_out
true
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_write.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
for cfg in 1 3 4; do
CONFIG=$cfg ncu --metrics $M --clock-control none -k regex:stencil_tma -s 5 -c 1 --csv ./benchmarks/micro/stencil_tma 40000 3 2>&1 | grep -v "^==" | tail -14
done
