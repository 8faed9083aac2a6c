#!/bin/bash
# GPU-box session: full ncu capture of the q8 kernels (one launch each, rough flow).
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
Q8_REPS=1 Q8_WARM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'q8_kernel' -c 5 -o $OUT/prof_q8 python tools/exp_q8_timing.py > $OUT/ncu_q8.log 2>&1
ls -la $OUT
