#!/bin/bash
# GPU-box session: compute-sanitizer over the kernels that changed in round 2 (shared-memory windows with barriers and
# atomics; packed-arithmetic 8-bit-frame kernels; packed-image warp)
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== racecheck (scatter windows)"; timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden_vectors or deterministic or packed_image or collapse or image_grad or batched" 2>&1 | tee $OUT/sanitizer_racecheck.log | grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY|hazard" | head -8
echo "== memcheck (8-bit-frame kernels, scatter)"; timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_q8_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "reference_fixture or vs_c_oracle or golden_vectors or large_flow or uint8_output or from_tables or windows_across" 2>&1 | tee $OUT/sanitizer_memcheck.log | grep -E "passed|failed|ERROR SUMMARY" | head -6
echo "== initcheck (smoke)"; timeout 600 compute-sanitizer --tool initcheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | tee $OUT/sanitizer_initcheck.log | grep -E "smoke ok|ERROR SUMMARY" | head -4
