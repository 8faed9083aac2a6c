#!/usr/bin/env python
"""GPU experiment: time ssm_upsample2x_nhwc on the U-Nets' decoder shapes (bf16, batch 2 at 1088x1920 output).
SSM_B200_LIB selects the library build (rows per thread / block size variants)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssm_b200 import unet_glue  # noqa: E402

dev = "cuda:0"
res = {"lib": os.path.basename(os.environ.get("SSM_B200_LIB", "default"))}
tot_ms, tot_bytes = 0.0, 0
for C, div in ((128, 2), (256, 4), (512, 8), (1024, 16), (1024, 32)):
    h, w = 1088 // div, 1920 // div
    x = torch.randn(2, C, h, w, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        y = unet_glue.upsample2x_cat([x])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = unet_glue.upsample2x_cat([x])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = 5 * x.numel() * 2
    res["C%d_1/%d" % (C, div)] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 0)}
    tot_ms += ms; tot_bytes += nbytes
    ref = torch.nn.functional.interpolate(x, size=(2 * h, 2 * w), mode="bilinear", align_corners=False)
    res["C%d_1/%d" % (C, div)]["max_diff_vs_aten"] = float((y.float() - ref.float()).abs().max())
res["total_ms"] = round(tot_ms, 4); res["total_gbs"] = round(tot_bytes / tot_ms / 1e6, 0)
print(json.dumps(res))
