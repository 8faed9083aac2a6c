#!/usr/bin/env python
"""GPU experiment: where does the in-situ time of the forward kernels go?

Compares, on the bench workload (16 pairs 1088x1920, N=7):
  A  each kernel alone, back to back (20 launches, events around each)
  B  the bench's pack -> flow_pack -> fuse loop with per-kernel events
  C  each kernel alone with a device sync + 30 ms idle before every launch (ncu-like isolation)
  D  plain device copy and fill bandwidth, burst (after idle) and sustained
Prints one JSON object; nothing here is a bench number.
"""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import synthetic  # noqa: E402

H, W, B, N = 1088, 1920, 16, 7
NPX = H * W
dev = torch.device("cuda:0")
img6 = synthetic.frames(B, H, W, seed=42, device=dev)
flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=dev)
out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=dev)
t = synthetic.timesteps(B, N, device=dev)
rgbx = ssm_b200.pack_frames(img6)
in16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=rgbx)
PACK_BYTES = (10 + 16 * N) * 4 * NPX * B
FUSE_BYTES = (6 + 12 * N) * 4 * NPX * B
RGBX_BYTES = 14 * 4 * NPX * B


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, reps=20, idle=0.0):
    out = []
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for _ in range(reps):
        if idle:
            torch.cuda.synchronize()
            time.sleep(idle)
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        out.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in out]
    return {"min": min(ms), "median": statistics.median(ms), "max": max(ms)}


res = {}
with torch.no_grad():
    k_rgbx = lambda: ssm_b200.pack_frames(img6)
    k_pack = lambda: ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=rgbx)
    k_pack_planar = lambda: ssm_b200.flow_pack(img6, flow4, t[:, :1].contiguous(), n_timesteps=1)
    k_fuse = lambda: ssm_b200.fuse(img6, in16, out5, t, packed=rgbx)
    for name, fn in (("rgbx", k_rgbx), ("flow_pack", k_pack), ("fuse", k_fuse)):
        res["A_alone_" + name] = timed(fn)
        res["C_isolated_" + name] = timed(fn, reps=8, idle=0.03)
    # B: bench-shaped loop
    evs = []
    for _ in range(3):
        r = k_rgbx(); x = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=r); ssm_b200.fuse(img6, x, out5, t, packed=r)
    torch.cuda.synchronize()
    for _ in range(20):
        e = [ev() for _ in range(4)]
        e[0].record(); r = k_rgbx(); e[1].record()
        x = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=r); e[2].record()
        ssm_b200.fuse(img6, x, out5, t, packed=r); e[3].record()
        evs.append(e)
    torch.cuda.synchronize()
    for i, name in enumerate(("rgbx", "flow_pack", "fuse")):
        res["B_loop_" + name] = {"median": statistics.median(e[i].elapsed_time(e[i + 1]) for e in evs)}
    # D: plain copy / fill
    big = torch.empty(PACK_BYTES // 4, dtype=torch.float32, device=dev)
    src = torch.empty(2 * 1024 ** 3, dtype=torch.float32, device=dev)
    dst = torch.empty_like(src)
    res["D_fill_16GB_sustained"] = timed(lambda: big.fill_(1.0))
    res["D_fill_16GB_isolated"] = timed(lambda: big.fill_(1.0), reps=8, idle=0.03)
    res["D_copy_8GB_sustained"] = timed(lambda: dst.copy_(src))
    res["D_copy_8GB_isolated"] = timed(lambda: dst.copy_(src), reps=8, idle=0.03)

gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9
summary = {
    "flow_pack_GBs": {k: gbs(PACK_BYTES, res[k + "_flow_pack"]["median"]) for k in ("A_alone", "B_loop", "C_isolated")},
    "fuse_GBs": {k: gbs(FUSE_BYTES, res[k + "_fuse"]["median"]) for k in ("A_alone", "B_loop", "C_isolated")},
    "rgbx_GBs": {k: gbs(RGBX_BYTES, res[k + "_rgbx"]["median"]) for k in ("A_alone", "B_loop", "C_isolated")},
    "fill_GBs": {k: gbs(PACK_BYTES, res["D_fill_16GB_" + k]["median"]) for k in ("sustained", "isolated")},
    "copy_GBs": {k: gbs(2 * src.numel() * 4, res["D_copy_8GB_" + k]["median"]) for k in ("sustained", "isolated")},
}
print(json.dumps({"raw_ms": res, "summary": summary}, indent=1))
