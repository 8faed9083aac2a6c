#!/bin/bash
# GPU-box sessions r04h / r04i: lazy flush of the scatter windows (SSM_SW_LAZY_FLUSH: the window of a frame survives
# consecutive timesteps while the tile's centre displacement stays within SSM_SW_DRIFT px of its origin) for several
# halo / drift pairs, next to the per-timestep flush; r04i also has max |G| recorded per timestep (no spills in the gather).
# usage: gpu_exp18.sh <tag> <variant>...   (variant = default or the <name> of tools/bin/libssm_<name>.so)
TAG=${1:-r04i}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  if [ $v = default ]; then unset SSM_B200_LIB; else export SSM_B200_LIB=$PWD/tools/bin/libssm_$v.so; fi
  echo "== $v"
  timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or image_grad or determin or collapse or large_flow" 2>&1 | tail -1 | tee $OUT/pytest_$v.log
  timeout 600 python tools/exp_bwd_timing.py 2>&1 | tail -1 > $OUT/bwd_timing_$v.json
  python - <<PY
import json
d=json.load(open("$OUT/bwd_timing_$v.json"))
for f in ("rough","smooth"):
    r=d[f]; print("$v", f, "fuse", round(r["fuse_bwd_gather_only"],2), round(r["fuse_bwd_with_image_grad"],2), round(r["ratio"],2), {k[:14]: round(x,2) for k,x in r["kernels_ms"].items()}, r["bit_identical_run_to_run"],
                  "| pack", round(r["flow_pack_bwd_gather_only"],2), round(r["flow_pack_bwd_with_image_grad"],2), {k[:14]: round(x,2) for k,x in r["flow_pack_kernels_ms"].items()})
PY
done
