#!/bin/bash
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest plumbing" ; timeout 600 python -m pytest tests/test_plumbing_gpu.py -x -q 2>&1 | tee $OUT/pytest_plumbing.log | tail -5
echo "== pipeline_step amp channels-last"
timeout 600 python tools/pipeline_step.py --amp --channels-last 2>&1 | tail -1 | tee -a $OUT/pipeline_step.jsonl
echo "== bench" ; timeout 900 python bench.py --no-train --no-e2e --no-cpu-baseline 2>&1 | tee $OUT/bench.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['kernels']); print(d['unet_layouts_variant'])"
