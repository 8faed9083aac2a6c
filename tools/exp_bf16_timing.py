#!/usr/bin/env python
"""GPU experiment: the 8-bit-frame kernels with bf16 storage of every tensor on the bench workload (one JSON line)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import q8, synthetic  # noqa: E402
from exp_q8_timing import timed  # noqa: E402

B, N, H, W = 16, 7, 1088, 1920
dev = torch.device("cuda:0")
x = synthetic.frames(2 * B, H, W, n_frames=1, seed=42, smooth=True, device=dev)
x = (x - x.amin()) / (x.amax() - x.amin())
images = (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
del x
lut = ssm_b200.normalisation_lut(device=dev)
planar, quads, norm, _ = q8.prepare(images, order="rgb", lut=lut, pad_values=lut[:, 0].tolist())
img_h = planar.view(B, 6, H, W).bfloat16()
flow_h = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=dev).bfloat16()
out5_h = synthetic.unet_out5(B, N, H, W, seed=44, device=dev).bfloat16()
t = synthetic.timesteps(B, N, device=dev)
in16 = torch.empty((B, N, 16, H, W), dtype=torch.bfloat16, device=dev)
out3 = torch.empty((B, N, 3, H, W), dtype=torch.bfloat16, device=dev)
with torch.no_grad():
    r = {"lib": os.environ.get("SSM_B200_LIB", "product"),
         "quads_from_u8": timed(lambda: q8.quads_from_u8(images, order="rgb")),
         "flow_pack_bf16": timed(lambda: q8.flow_pack(img_h, quads, flow_h, t, norm, n_timesteps=N, out=in16)),
         "fuse_bf16": timed(lambda: q8.fuse_from_flow(quads, flow_h, out5_h, t, norm, out=out3))}
r["path_ms"] = r["quads_from_u8"] + r["flow_pack_bf16"] + r["fuse_bf16"]
r["path_frac_of_6551"] = ((10 + 16 * N) + (10 + 8 * N)) * 2 * H * W * B / (r["path_ms"] * 1e-3) / 1e9 / 6551.0
print(json.dumps(r))
