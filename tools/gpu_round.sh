#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
if [ "$2" != "quick" ]; then
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tee $OUT/pytest_gpu.log | tail -15
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tee $OUT/smoke.log | tail -3
fi
echo "== bench" ; timeout 1200 python bench.py --steps 20 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.log | tail -1 | cut -c1-3000; tail -5 $OUT/bench.err
if [ "$2" != "quick" ]; then
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee $OUT/bench_reference.log | tail -1 | cut -c1-400
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(warp_|flow_pack_|fuse_|scatter_|absmax_|pack_frames|quads_)' -c 60 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-variants --no-configs --no-reference-gpu > $OUT/ncu_launches.log 2>&1
tail -12 $OUT/launches.csv
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'flow_pack_fwd|fuse_fwd|pack_frames|quads_from' -s 12 -c 6 \
    -o $OUT/prof_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-variants --no-configs --no-reference-gpu > $OUT/ncu_full.log 2>&1
fi
ls -la $OUT
