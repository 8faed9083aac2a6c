#!/bin/bash
# GPU-box session r02k: image-gradient scatter schemes (tools/exp_scatter.cu) + the new parity tests
TAG=${1:-r02k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 tools/bin/exp_scatter 4 2>&1 | tee $OUT/exp_scatter.jsonl
echo "== pytest (new tests)"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "large_flow or patch_reference or device_side_t or packed_image" 2>&1 | tee $OUT/pytest_new.log | tail -8
