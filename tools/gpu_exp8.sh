#!/bin/bash
# GPU-box session: timed run of the gather microbenchmark (events, not under a profiler) + L1 counters (one launch each).
TAG=${1:-r01p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 tools/bin/exp_gather > $OUT/exp_gather.jsonl 2>&1
grep -E "mixlin|texw0|lanepair|\"rgbx16\"" $OUT/exp_gather.jsonl | cut -c1-120
bash tools/gpu_exp7.sh $TAG > /dev/null 2>&1
wc -l $OUT/exp_gather_counters.csv
