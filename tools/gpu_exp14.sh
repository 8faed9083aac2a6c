#!/bin/bash
# GPU-box session: whole GPU suite + bench line of the current tree
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tee $OUT/pytest_gpu.log | tail -8
echo "== bench" ; timeout 1200 python bench.py --steps 20 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.log | tail -1 | cut -c1-300; tail -5 $OUT/bench.err
