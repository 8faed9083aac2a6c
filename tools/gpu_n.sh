#!/bin/bash
# multi-GPU session: both bench arms under torchrun, as the driver launches them.  usage: bash tools/gpu_n.sh N tag
N=${1:-2}
TAG=${2:-r02n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia-smi.txt 2>&1
echo "== bench N=$N"
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) 2>$OUT/bench.err | tee $OUT/bench.log | tail -1 | cut -c1-600
tail -4 $OUT/bench.err
echo "== reference arm N=$N"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) 2>$OUT/bench_reference.err | tee $OUT/bench_reference.log | tail -1 | cut -c1-400
tail -4 $OUT/bench_reference.err
