#!/bin/bash
# GPU-box session r04f: bf16 storage with the direct sampling coordinate x + flow (SSM_Q8_BF16_DIRECT_COORD), with and
# without the difference-form interpolation, against the shipped build; bf16 parity tests on the variant.
TAG=${1:-r04f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default bf16_direct bf16_direct_diff; do
  if [ $v = default ]; then unset SSM_B200_LIB; else export SSM_B200_LIB=$PWD/tools/bin/libssm_$v.so; fi
  echo "== $v"
  if [ $v != default ]; then timeout 600 python -m pytest tests/test_q8_gpu.py -m gpu -x -q -k "bf16 or storage" 2>&1 | tail -2 | tee $OUT/pytest_q8_$v.log; fi
  Q8_REPS=20 timeout 300 python tools/exp_bf16_timing.py 2>&1 | tail -1 | tee $OUT/bf16_$v.json
done
