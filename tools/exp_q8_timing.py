#!/usr/bin/env python
"""GPU experiment: the 8-bit-frame kernels against the fp32 RGBx kernels on the bench workload
(16 pairs 1088x1920 x 7 timesteps), CUDA events, 10 repetitions after 3 warm-ups.  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import q8, synthetic  # noqa: E402

B, N, H, W = 16, 7, 1088, 1920
NPX = H * W
dev = torch.device("cuda:0")


def timed(fn, reps=int(os.environ.get("Q8_REPS", "10")), warm=int(os.environ.get("Q8_WARM", "3"))):
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


def main():
    g = torch.Generator().manual_seed(3)
    x = synthetic.frames(2 * B, H, W, n_frames=1, seed=42, smooth=True, device=dev)
    x = (x - x.amin()) / (x.amax() - x.amin())
    images = (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
    del x
    lut = ssm_b200.normalisation_lut(device=dev)
    pads = lut[:, 0].tolist()
    planar, quads, norm, _ = q8.prepare(images, order="rgb", lut=lut, pad_values=pads)
    img6 = planar.view(B, 6, H, W)
    res = {}
    for name, div in (("rough", 8), ("smooth", 64)):
        c = torch.randn((B, 4, H // div, W // div), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 20.0
        flow4 = torch.nn.functional.interpolate(c, size=(H, W), mode="bilinear", align_corners=False).contiguous()
        out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=dev)
        t = synthetic.timesteps(B, N, device=dev)
        in16 = torch.empty((B, N, 16, H, W), device=dev)
        out3 = torch.empty((B, N, 3, H, W), device=dev)
        rgbx = torch.empty((B, 2, H, W, 4), device=dev)
        out_u8 = torch.empty((B, N, H, W, 3), dtype=torch.uint8, device=dev)
        nhwc = torch.empty((B, N, H, W, 16), dtype=torch.bfloat16, device=dev).permute(0, 1, 4, 2, 3)
        y16 = out5.bfloat16()
        with torch.no_grad():
            r = {
                "quads_from_u8": timed(lambda: q8.quads_from_u8(images, order="rgb")),
                "frames_from_u8_planar": timed(lambda: ssm_b200.frames_from_u8(images, order="rgb", lut=lut, pad_values=pads)),
                "q8_flow_pack": timed(lambda: q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N, out=in16)),
                "q8_flow_pack_from_tables": timed(lambda: q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N, out=in16, lut=lut)),
                "q8_fuse": timed(lambda: q8.fuse_from_flow(quads, flow4, out5, t, norm, out=out3)),
                "q8_fuse_bf16_out5": timed(lambda: q8.fuse_from_flow(quads, flow4, y16, t, norm, out=out3)),
                "q8_fuse_to_u8": timed(lambda: q8.fuse_from_flow_to_u8(quads, flow4, out5, t, norm, out=out_u8)),
                "q8_flow_pack_nhwc_bf16": timed(lambda: q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N, out=nhwc,
                                                                       channels_last_dtype=torch.bfloat16)),
                "fp32_pack_frames": timed(lambda: ssm_b200.pack_frames(img6, out=rgbx)),
                "fp32_flow_pack": timed(lambda: ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=rgbx, out=in16)),
                "fp32_fuse": timed(lambda: ssm_b200.fuse_from_flow(img6, flow4, out5, t, packed=rgbx, out=out3)),
            }
        k1_bytes = (10 + 16 * N) * 4 * NPX * B
        k2_bytes = (10 + 8 * N) * 4 * NPX * B
        r["q8_flow_pack_gbs"] = k1_bytes / r["q8_flow_pack"] / 1e6
        r["q8_fuse_gbs"] = (4 * 4 + 2 * 4 + 8 * 4 * N) * NPX * B / r["q8_fuse"] / 1e6
        r["fp32_flow_pack_gbs"] = k1_bytes / r["fp32_flow_pack"] / 1e6
        r["fp32_fuse_gbs"] = k2_bytes / r["fp32_fuse"] / 1e6
        res[name] = r
        del flow4, out5, in16, out3, rgbx, out_u8, nhwc, y16
    print(json.dumps(res))


if __name__ == "__main__":
    main()
