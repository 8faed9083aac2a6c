#!/bin/bash
# backward kernels: launch list + full capture (training case: no image gradients)
OUT=gpurun_out/${1:-exp3}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bwd_launches.csv \
   python tools/profile_kernels.py --bwd --reps 2 > $OUT/bwd_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bwd_c3_launches.csv \
   python tools/profile_kernels.py --bwd --reps 2 --pairs 64 --timesteps 1 --height 352 --width 352 > $OUT/bwd_c3_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fuse_bwd|flow_pack_bwd' -c 2 \
   -o $OUT/prof_bwd python tools/profile_kernels.py --bwd --reps 1 --pairs 4 > $OUT/ncu_bwd_full.log 2>&1
ls -la $OUT
