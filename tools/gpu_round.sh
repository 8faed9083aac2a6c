#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tee $OUT/pytest_gpu.log | tail -40
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $OUT/smoke.log | tail -5
echo "== bench" ; timeout 900 python bench.py 2>&1 | tee $OUT/bench.log | tail -3
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee $OUT/bench_reference.log | tail -2
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(warp_|flow_pack_|fuse_|scatter_|absmax_|pack_frames)' -c 40 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-variants > $OUT/ncu_launches.log 2>&1
tail -12 $OUT/launches.csv
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'flow_pack_fwd|fuse_fwd|pack_frames' -s 9 -c 3 \
    -o $OUT/prof_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-variants > $OUT/ncu_full.log 2>&1
ls -la $OUT
