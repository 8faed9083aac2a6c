#!/bin/bash
# GPU-box session r04g: what recording max |G| (one L2 load + atomicMax per warp on ONE address) costs the gather kernel
# of the image-gradient backward: shipped build vs a timing-only build without the recording.
TAG=${1:-r04g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default noabsmax; do
  if [ $v = default ]; then unset SSM_B200_LIB; else export SSM_B200_LIB=$PWD/tools/bin/libssm_$v.so; fi
  echo "== $v"
  timeout 600 python tools/exp_bwd_timing.py 2>&1 | tail -1 | tee $OUT/bwd_timing_$v.json | cut -c1-900
done
