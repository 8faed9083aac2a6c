#!/bin/bash
# GPU-box session r01n: L1 counters of the gather microbenchmark variants (one launch each, rough flow).
TAG=${1:-r01n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_' --csv --log-file $OUT/exp_gather_counters.csv tools/bin/exp_gather once > $OUT/exp_gather_once.log 2>&1
tail -3 $OUT/exp_gather_once.log
wc -l $OUT/exp_gather_counters.csv
