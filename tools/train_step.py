#!/usr/bin/env python
"""Data-parallel training step of the full model on synthetic crops (BASELINE.json configs[2]:
352x352 crops, global batch 64, warp backward), one process per GPU:

    python tools/train_step.py [--steps 10] [--warmup 3] [--global-batch 64] [--size 352] [--n-frames 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py ...

The two U-Nets are stock torch/cuDNN modules with random-init weights (seed 42) wrapped in
DistributedDataParallel: their parameter gradients are the ONLY thing that crosses NVLink (NCCL
all-reduce, overlapped with backward by DDP's buckets).  The synthesis path between and after them
(compute_inputs, compute_output_image + loss front-end, and their backward) is this repo's kernels;
it has no parameters and no collective.  Prints one JSON line (not the bench.py headline):
samples/s over all ranks, ms/step (max over ranks, CUDA events), and the share of the step spent in
this repo's kernels, measured in a second pass with events around every call of the path.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import synthetic  # noqa: E402
from ssm_b200.superslomo_r import FullModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--global-batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=352)
    ap.add_argument("--n-frames", type=int, default=2, help="2 = SuperSloMo original (one window), 4 = SSMR windows")
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the U-Nets (the path stays fp32)")
    ap.add_argument("--channels-last", action="store_true", help="NHWC U-Nets (cuDNN tensor-core kernels)")
    ap.add_argument("--bottleneck", default="CONV", choices=["CONV", "CLSTM", "CGRU"],
                    help="U-Net bottleneck: CONV = superslomo_original.ini, CLSTM = superslomo_recurrent.ini (use --n-frames 4)")
    ap.add_argument("--profile", default=None, help="write the kernel table of one step (torch.profiler) to this file")
    a = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert a.global_batch % world == 0
    B, S, T = a.global_batch // world, a.size, a.n_frames

    torch.manual_seed(42)                       # same initial weights on every rank (configs/*.ini [SEED])
    import configparser
    cfg = configparser.RawConfigParser()
    cfg.read_string("[STAGE1]\nBOTTLENECK=%s\n[STAGE2]\nBOTTLENECK=%s\nCROSS_SKIP=TRUE\n" % (a.bottleneck, a.bottleneck))
    model = FullModel(cfg=cfg, loss=ssm_b200.losses.SSMLosses(lambda_r=60.0, lambda_p=0.0, lambda_w=10.0)).to(dev)
    if a.channels_last:
        model.stage1_model.set_channels_last()
        model.stage2_model.set_channels_last()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    n_params = sum(p.numel() for p in model.parameters())

    frames = synthetic.frames(B, S, S, n_frames=T, seed=100 + rank, device=dev).view(B, T, 3, S, S)
    targets = synthetic.frames(B, S, S, n_frames=T - 1, seed=200 + rank, device=dev).view(B, T - 1, 3, S, S)
    t = synthetic.random_timesteps(B, T - 1, seed=300 + rank).to(dev).view(B, T - 1, 1, 1, 1)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            _, losses = net(frames, t, target_images=targets, inference_mode=False)
        loss = losses[:, 0].float().mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()

    # share of the step spent in this repo's kernels: events around every C-ABI call (fwd and bwd)
    spans = []
    lib = ssm_b200._abi.lib()
    names = ["ssm_pack_frames", "ssm_flow_pack_fwd", "ssm_flow_pack_bwd", "ssm_fuse_loss_fwd", "ssm_fuse_loss_bwd",
             "ssm_fuse_flow_fwd", "ssm_fuse_flow_bwd", "ssm_warp_fwd", "ssm_warp_bwd"]
    originals = {n: getattr(lib, n) for n in names}

    def timed(n, fn):
        def call(*args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            spans.append((n, s, e))
            return rc
        return call

    class _Shim:          # ctypes function objects cannot be replaced on the CDLL: shim the attribute lookup
        def __getattr__(self, n):
            return timed(n, originals[n]) if n in originals else getattr(lib, n)

    ssm_b200._abi._lib = _Shim()
    step()
    torch.cuda.synchronize()
    ssm_b200._abi._lib = lib
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step()
            torch.cuda.synchronize()
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=50, max_name_column_width=90))
    path_ms = {}
    for n, s, e in spans:
        path_ms[n] = path_ms.get(n, 0.0) + s.elapsed_time(e)

    if rank == 0:
        print(json.dumps({
            "what": "data-parallel training step (synthetic crops, random-init U-Nets)", "n_gpus": world,
            "global_batch": a.global_batch, "per_gpu_batch": B, "crop": S, "n_frames": T, "bottleneck": a.bottleneck, "amp_bf16_unets": a.amp, "channels_last_unets": a.channels_last,
            "ms_per_step": ms, "samples_per_s": a.global_batch / (ms * 1e-3), "loss": float(loss.detach()),
            "unet_parameters": n_params, "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
            "path_kernels_ms": path_ms, "path_ms_total": sum(path_ms.values()),
            "path_share_of_step": sum(path_ms.values()) / ms}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
