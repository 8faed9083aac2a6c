#!/usr/bin/env python
"""Launch the forward (and optionally backward) kernels a few times on the bench workload so that
ncu can capture them:  python tools/profile_kernels.py [--flow rough|smooth] [--mode cpu|cuda]
[--bwd] [--pairs B] [--reps R].  Prints event timings as JSON (not a bench number under ncu)."""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--flow", default="rough")
ap.add_argument("--mode", default="cpu")
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--img-grad", action="store_true")
ap.add_argument("--pairs", type=int, default=16)
ap.add_argument("--timesteps", type=int, default=7)
ap.add_argument("--height", type=int, default=1088)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--from-flow", action="store_true", help="fuse_from_flow (estimated flows recomputed in-kernel)")
a = ap.parse_args()

H, W, B, N = a.height, a.width, a.pairs, a.timesteps
NPX = H * W
dev = torch.device("cuda:0")
dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
img6 = synthetic.frames(B, H, W, seed=42, device=dev).to(dt)
if a.flow == "rough":
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=dev)
else:
    g = torch.Generator(device=dev).manual_seed(7)
    flow4 = torch.nn.functional.interpolate(torch.randn((B, 4, max(H // 64, 2), max(W // 64, 2)), device=dev, generator=g) * 20.0,
                                            size=(H, W), mode="bilinear", align_corners=False).contiguous()
flow4 = flow4.to(dt)
out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=dev).to(dt)
t = synthetic.timesteps(B, N, device=dev)
ssm_b200.set_coord_mode(a.mode)
esz = 4 if a.dtype == "f32" else 2


def ev():
    return torch.cuda.Event(enable_timing=True)


times = {"rgbx": [], "flow_pack": [], "fuse": [], "fuse_bwd": [], "flow_pack_bwd": []}
for _ in range(a.reps):
    e = [ev() for _ in range(7)]
    if a.bwd:
        f = flow4.clone().requires_grad_(True)
        y = out5.clone().requires_grad_(True)
        im = img6.clone().requires_grad_(a.img_grad)
    else:
        f, y, im = flow4, out5, img6
    with torch.set_grad_enabled(a.bwd):
        e[0].record()
        rgbx = ssm_b200.pack_frames(im)
        e[1].record()
        in16 = ssm_b200.flow_pack(im, f, t, n_timesteps=N, packed=rgbx)
        e[2].record()
        fr = ssm_b200.fuse_from_flow(im, f, y, t, packed=rgbx) if a.from_flow else ssm_b200.fuse(im, in16, y, t, packed=rgbx)
        e[3].record()
    if a.bwd:
        g3 = torch.randn_like(fr)
        torch.cuda.synchronize()
        e[3].record()
        src = f if a.from_flow else in16
        (gin16, gy) = torch.autograd.grad(fr, (src, y), g3, retain_graph=True) if not a.img_grad else torch.autograd.grad(fr, (src, y, im), g3, retain_graph=True)[:2]
        e[4].record()
        if a.from_flow:   # stand-in for the stage-2 U-Net's gradient of its 16-channel input
            gin16 = torch.randn_like(in16)
            torch.cuda.synchronize()
        e[6].record()
        torch.autograd.grad(in16, (f, im) if a.img_grad else (f,), gin16)
        e[5].record()
    torch.cuda.synchronize()
    times["rgbx"].append(e[0].elapsed_time(e[1]))
    times["flow_pack"].append(e[1].elapsed_time(e[2]))
    times["fuse"].append(e[2].elapsed_time(e[3]) if not a.bwd else float("nan"))
    if a.bwd:
        times["fuse_bwd"].append(e[3].elapsed_time(e[4]))
        times["flow_pack_bwd"].append(e[6].elapsed_time(e[5]))
    del in16, fr, rgbx

algo = {"rgbx": 14 * esz * NPX * B, "flow_pack": (10 + 16 * N) * esz * NPX * B, "fuse": ((10 + 8 * N) if a.from_flow else (6 + 12 * N)) * esz * NPX * B,
        "fuse_bwd": ((10 + 13 * N) if a.from_flow else (6 + 21 * N)) * esz * NPX * B, "flow_pack_bwd": (14 + 10 * N) * esz * NPX * B}
out = {"args": vars(a)}
for k, v in times.items():
    v = [x for x in v[1:] if x == x]
    if v:
        ms = statistics.median(v)
        out[k] = {"ms": ms, "algorithmic_GBs": algo[k] / (ms * 1e-3) / 1e9}
print(json.dumps(out))
