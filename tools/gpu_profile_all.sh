#!/bin/bash
# GPU-box session: `ncu --set full` over one launch of every kernel family (tools/profile_all.py); the raw-page CSV and a
# per-kernel summary come back, the (large) report stays on the box except for the scatter kernels' own capture.
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
K='regex:^(warp_|flow_pack_|fuse_|scatter_|absmax_|pack_frames|pack_image|quads_|frames_|loss_)'
timeout 1500 ncu --set full --clock-control none -k "$K" -o /tmp/prof_all python tools/profile_all.py c2 c3 > $OUT/ncu_all.log 2>&1
tail -3 $OUT/ncu_all.log
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > $OUT/ncu_all_raw.csv 2>/dev/null
ls -la /tmp/prof_all.ncu-rep $OUT
python tools/ncu_summary.py $OUT/ncu_all_raw.csv > $OUT/ncu_all_summary.txt 2>&1
grep -c "^== " $OUT/ncu_all_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fuse_scatter_kernel|warp_scatter_win' -c 2 -o $OUT/prof_scatter python tools/profile_all.py c2 c3 > $OUT/ncu_scatter.log 2>&1
ls -la $OUT
