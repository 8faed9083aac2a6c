#!/bin/bash
# GPU-box session r01j: plumbing tests, pipeline step with/without the U-Net layouts, bench with the new variant.
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi.txt 2>&1
echo "== pytest plumbing" ; timeout 600 python -m pytest tests/test_plumbing_gpu.py tests/test_model_loop_gpu.py -x -q 2>&1 | tee $OUT/pytest_plumbing.log | tail -15
for extra in "--generic-plumbing" ""; do
  echo "== pipeline_step amp channels-last $extra"
  timeout 600 python tools/pipeline_step.py --amp --channels-last $extra 2>&1 | tail -1 | tee -a $OUT/pipeline_step.jsonl
done
echo "== bench" ; timeout 900 python bench.py --no-train 2>&1 | tee $OUT/bench.log | tail -1 | cut -c1-3000
