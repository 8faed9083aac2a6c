#!/bin/bash
# GPU session: in-situ vs isolated timings, backward timings (C2 and C3 shapes), bf16 storage timings
OUT=gpurun_out/${1:-exp2}
mkdir -p $OUT
timeout 300 python tools/exp_timing.py > $OUT/exp_timing.json 2> $OUT/exp_timing.err
for args in "--bwd" "--bwd --img-grad" "--bwd --pairs 64 --timesteps 1 --height 352 --width 352 --reps 8" \
            "--bwd --img-grad --pairs 64 --timesteps 1 --height 352 --width 352 --reps 8" \
            "--dtype bf16" "--dtype bf16 --bwd" "--flow smooth" "--mode cuda" \
            "--pairs 1 --timesteps 31 --height 2176 --width 3840" "--pairs 1 --height 736 --width 1280"; do
  timeout 300 python tools/profile_kernels.py $args 2>> $OUT/profile_kernels.err | tee -a $OUT/profile_kernels.jsonl
done
tail -5 $OUT/exp_timing.json
