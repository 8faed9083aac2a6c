#!/bin/bash
# Short GPU-box session: parity tests, smoke (also under compute-sanitizer memcheck), bench line.
# Usage (from the repo root on the GPU box):  bash tools/gpu_verify.sh [tag]
TAG=${1:-verify}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tee $OUT/pytest_gpu.log | tail -15
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $OUT/smoke.log | tail -3
echo "== smoke under compute-sanitizer (memcheck)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > $OUT/sanitizer_memcheck.log 2>&1
echo "exit $?" | tee -a $OUT/sanitizer_memcheck.log ; tail -4 $OUT/sanitizer_memcheck.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tee $OUT/bench.log | tail -2
ls -la $OUT
