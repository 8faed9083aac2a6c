#!/usr/bin/env python
"""GPU experiment: the backward with image gradients (deterministic segmented scatter) on the bench workload
(16 pairs 1088x1920 x 7 timesteps), rough (1/8-resolution control grid) and smooth (1/64) flow fields: whole-backward
time next to the gather-only backward, and the per-kernel split from torch.profiler.  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import synthetic  # noqa: E402

B, N, H, W = 16, 7, 1088, 1920
dev = torch.device("cuda:0")


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(reps):
        fn()
    e[1].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) / reps


def kernel_split(fn):
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    out = {}
    for ev in prof.key_averages():
        name = ev.key.split("<")[0].replace("void ssm::", "").replace("void ", "")
        t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
        if t > 0:
            out[name] = out.get(name, 0.0) + t / 1e3
    return dict(sorted(out.items(), key=lambda kv: -kv[1])[:8])


def main():
    img6 = synthetic.frames(B, H, W, seed=42, smooth=True, device=dev)
    out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=dev)
    t = synthetic.timesteps(B, N, device=dev)
    res = {}
    for name, div in (("rough", 8), ("smooth", 64)):
        c = torch.randn((B, 4, H // div, W // div), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 20.0
        flow4 = torch.nn.functional.interpolate(c, size=(H, W), mode="bilinear", align_corners=False).contiguous()
        ig, fg, yg = img6.clone().requires_grad_(True), flow4.clone().requires_grad_(True), out5.clone().requires_grad_(True)
        rgbx = ssm_b200.pack_frames(img6)
        frames = ssm_b200.fuse_from_flow(ig, fg, yg, t, packed=rgbx)
        frames_no = ssm_b200.fuse_from_flow(img6, fg, yg, t, packed=rgbx)
        g3 = torch.randn_like(frames)
        r = {"fuse_bwd_gather_only": timed(lambda: torch.autograd.grad(frames_no, (fg, yg), g3, retain_graph=True)),
             "fuse_bwd_with_image_grad": timed(lambda: torch.autograd.grad(frames, (ig, fg, yg), g3, retain_graph=True))}
        r["ratio"] = r["fuse_bwd_with_image_grad"] / r["fuse_bwd_gather_only"]
        r["kernels_ms"] = kernel_split(lambda: torch.autograd.grad(frames, (ig, fg, yg), g3, retain_graph=True))
        # run-to-run determinism of the image gradient at full size
        a = torch.autograd.grad(frames, (ig,), g3, retain_graph=True)[0]
        b = torch.autograd.grad(frames, (ig,), g3, retain_graph=True)[0]
        r["bit_identical_run_to_run"] = bool(torch.equal(a, b))
        del frames, frames_no, g3, a, b, ig, fg, yg
        xw, fw = img6[:, 0:3].contiguous().requires_grad_(True), (0.5 * flow4[:, 0:2]).contiguous().requires_grad_(True)
        yw = ssm_b200.warp(xw, fw)
        gw = torch.randn_like(yw)
        r["warp_bwd_flow_and_image"] = timed(lambda: torch.autograd.grad(yw, (xw, fw), gw, retain_graph=True))
        r["warp_kernels_ms"] = kernel_split(lambda: torch.autograd.grad(yw, (xw, fw), gw, retain_graph=True))
        del yw, gw, xw, fw
        # compute_inputs backward: flow gradient only / with the image gradient
        ig, fg = img6.clone().requires_grad_(True), flow4.clone().requires_grad_(True)
        in16 = ssm_b200.flow_pack(ig, fg, t, n_timesteps=N, packed=rgbx)
        in16_no = ssm_b200.flow_pack(img6, fg, t, n_timesteps=N, packed=rgbx)
        g16 = torch.randn_like(in16)
        r["flow_pack_bwd_gather_only"] = timed(lambda: torch.autograd.grad(in16_no, (fg,), g16, retain_graph=True))
        r["flow_pack_bwd_with_image_grad"] = timed(lambda: torch.autograd.grad(in16, (ig, fg), g16, retain_graph=True))
        r["flow_pack_kernels_ms"] = kernel_split(lambda: torch.autograd.grad(in16, (ig, fg), g16, retain_graph=True))
        del in16, in16_no, g16, ig, fg
        res[name] = r
        del flow4, rgbx
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
