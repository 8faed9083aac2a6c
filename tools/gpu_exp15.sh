#!/bin/bash
# GPU-box session r04e: compute_output_image with the gathers of timestep n+1 issued before the interpolation of timestep n
# (SSM_Q8_FUSE_PIPE) at 2 and 3 CTAs/SM, against the shipped loop at 3 and 2 CTAs/SM.
TAG=${1:-r04e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default pipe2 pipe3 nopipe2; do
  if [ $v = default ]; then unset SSM_B200_LIB; else export SSM_B200_LIB=$PWD/tools/bin/libssm_$v.so; fi
  echo "== $v"
  if [ $v = pipe2 ]; then timeout 600 python -m pytest tests/test_q8_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee $OUT/pytest_q8_$v.log; fi
  Q8_REPS=20 timeout 300 python tools/exp_q8_timing.py 2>&1 | tail -1 > $OUT/q8_timing_$v.json
  python - <<PY
import json
d=json.load(open("$OUT/q8_timing_$v.json"))
for f in ("rough","smooth"):
    r=d[f]; print("$v", f, {k: round(r[k],3) for k in ("q8_flow_pack","q8_fuse","q8_fuse_bf16_out5","q8_fuse_to_u8")})
PY
done
