#!/bin/bash
mkdir -p gpurun_out/r01ups
for v in 1_256 2_256 4_256 4_128 8_128; do
  SSM_B200_LIB=$PWD/tools/bin/libssm_ups_$v.so timeout 120 python tools/exp_upsample.py 2>&1 | tail -1 | tee -a gpurun_out/r01ups/exp_upsample.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['lib'], d['total_ms'], d['total_gbs'], d['C128_1/2'])"
done
