#!/bin/bash
# GPU-box session r02h: packed-fp32 (two pixels per instruction) q8 kernels -- parity tests, then timing of the default
# occupancy (compute_inputs 3 CTAs/SM, compute_output_image 4) and the alternative build (4 / 3).
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia-smi.txt 2>&1
echo "== pytest q8" ; timeout 900 python -m pytest tests/test_q8_gpu.py -m gpu -x -q 2>&1 | tee $OUT/pytest_q8.log | tail -8
echo "== timing default" ; timeout 300 python tools/exp_q8_timing.py 2>&1 | tee $OUT/q8_timing_default.json | cut -c1-1200
echo "== timing alt" ; SSM_B200_LIB=$PWD/tools/bin/libssm_q8_alt.so timeout 300 python tools/exp_q8_timing.py 2>&1 | tee $OUT/q8_timing_alt.json | cut -c1-1200
