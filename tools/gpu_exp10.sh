#!/bin/bash
# GPU-box session r02a: round-2 gather microbenchmark (smaller texels, smem staging): timed run + L1 counters.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia-smi.txt 2>&1
timeout 600 tools/bin/exp_gather2 > $OUT/exp_gather2.jsonl 2>&1
cat $OUT/exp_gather2.jsonl | cut -c1-140
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_' --csv --log-file $OUT/exp_gather2_counters.csv tools/bin/exp_gather2 once > $OUT/exp_gather2_once.log 2>&1
tail -3 $OUT/exp_gather2_once.log
wc -l $OUT/exp_gather2_counters.csv
