#!/bin/bash
# backward + fused-loss kernels after the rewrite: launch list and full capture
OUT=gpurun_out/${1:-exp4}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(warp_|flow_pack_|fuse_|scatter_|absmax_|pack_frames|loss_reduce|frames_)' --csv --log-file $OUT/bwd_launches.csv \
   python tools/profile_kernels.py --bwd --from-flow --reps 2 > $OUT/bwd_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fuse_bwd|flow_pack_bwd' -c 2 \
   -o $OUT/prof_bwd python tools/profile_kernels.py --bwd --from-flow --reps 1 --pairs 4 > $OUT/ncu_bwd_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(warp_|flow_pack_|fuse_|scatter_|absmax_|pack_frames|loss_reduce|frames_)' --csv --log-file $OUT/train_launches.csv \
   python tools/train_step.py --steps 1 --warmup 1 > $OUT/train_launches.log 2>&1
tail -3 $OUT/train_launches.log
ls -la $OUT
