#!/usr/bin/env python
"""One launch of every kernel family of the library at BASELINE sizes, for `ncu --set full`:
C2 size (16 pairs 1088x1920 x 7 timesteps) for the path kernels, forward and backward, fp32 / bf16 storage / uint8 frames;
C3 size (64 x 352 x 352, N = 1) for the fused-loss kernels and the stand-alone warp.  Not a bench number."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import q8, synthetic  # noqa: E402

dev = torch.device("cuda:0")
PAIRS = int(os.environ.get("PROFILE_PAIRS", "16"))


def c2():
    B, N, H, W = PAIRS, 7, 1088, 1920
    x = synthetic.frames(2 * B, H - 8, W, n_frames=1, seed=42, smooth=True, device=dev)
    x = (x - x.amin()) / (x.amax() - x.amin())
    images = (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
    del x
    lut = ssm_b200.normalisation_lut(device=dev)
    planar, quads, norm, _ = q8.prepare(images, order="rgb", lut=lut, pad_values=lut[:, 0].tolist())
    img6 = planar.view(B, 6, H, W)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=dev)
    out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=dev)
    t = synthetic.timesteps(B, N, device=dev)
    with torch.no_grad():
        # uint8 frames (headline)
        q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N)
        q8.fuse_from_flow(quads, flow4, out5, t, norm)
        q8.fuse_from_flow(quads, flow4, out5.bfloat16(), t, norm)
        out_u8 = torch.empty((B, N, H - 8, W, 3), dtype=torch.uint8, device=dev)
        q8.fuse_from_flow_to_u8(quads, flow4, out5, t, norm, crop=(4, 0, H - 8, W), out=out_u8)
        del out_u8
        nhwc = torch.empty((B, N, H, W, 16), dtype=torch.bfloat16, device=dev).permute(0, 1, 4, 2, 3)
        q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N, out=nhwc, channels_last_dtype=torch.bfloat16)
        del nhwc
        # bf16 storage of every tensor, uint8 frames
        ih, fh, yh = img6.bfloat16(), flow4.bfloat16(), out5.bfloat16()
        q8.flow_pack(ih, quads, fh, t, norm, n_timesteps=N)
        q8.fuse_from_flow(quads, fh, yh, t, norm)
        # bf16 storage, bf16 RGBx gathers
        r = ssm_b200.pack_frames(ih)
        ssm_b200.flow_pack(ih, fh, t, n_timesteps=N, packed=r)
        ssm_b200.fuse_from_flow(ih, fh, yh, t, packed=r)
        del ih, fh, yh, r
        # fp32 frames
        rgbx = ssm_b200.pack_frames(img6)
        ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=rgbx)
        ssm_b200.fuse_from_flow(img6, flow4, out5, t, packed=rgbx)
        frames = ssm_b200.fuse_from_flow(img6, flow4, out5, t, packed=rgbx)
        ssm_b200.frames_to_u8(frames[:, 0].contiguous(), top=4, left=0, h_out=H - 8, w_out=W)
        del frames
    torch.cuda.empty_cache()
    # backward: gather-only, then with image gradients (segmented scatter)
    fg, yg = flow4.clone().requires_grad_(True), out5.clone().requires_grad_(True)
    in16 = ssm_b200.flow_pack(img6, fg, t, n_timesteps=N, packed=rgbx)
    fr = ssm_b200.fuse_from_flow(img6, fg, yg, t, packed=rgbx)
    g3 = torch.randn_like(fr)
    torch.autograd.grad(fr, (fg, yg), g3)
    del fr
    g16 = torch.randn_like(in16)
    torch.autograd.grad(in16, (fg,), g16)
    del in16
    torch.cuda.empty_cache()
    ig = img6.clone().requires_grad_(True)
    fr = ssm_b200.fuse_from_flow(ig, fg, yg, t, packed=rgbx)
    torch.autograd.grad(fr, (ig, fg, yg), g3)
    del fr, g3
    in16 = ssm_b200.flow_pack(ig, fg, t, n_timesteps=N, packed=rgbx)
    torch.autograd.grad(in16, (ig, fg), g16)
    del in16, g16
    torch.cuda.empty_cache()


def c3():
    B, H, W = 64, 352, 352
    img6 = synthetic.frames(B, H, W, seed=1, device=dev)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=4.0, seed=2, device=dev).requires_grad_(True)
    out5 = synthetic.unet_out5(B, 1, H, W, seed=3, device=dev).requires_grad_(True)
    target = synthetic.frames(B, H, W, n_frames=1, seed=4, device=dev).view(B, 1, 3, H, W)
    t = synthetic.random_timesteps(B, 1, seed=5).to(dev)
    frames, sums = ssm_b200.fuse_loss(img6, flow4, out5, target, t)
    (sums.sum() + frames.sum()).backward()
    x, f = img6[:, 0:3].contiguous().requires_grad_(True), flow4.detach()[:, 0:2].contiguous().requires_grad_(True)
    y = ssm_b200.warp(x, f)
    y.sum().backward()
    pk = ssm_b200.pack_image(x)
    y = ssm_b200.warp(x, f, packed=pk)
    torch.autograd.grad(y, (f,), torch.ones_like(y))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3"]
    if "c2" in which:
        c2()
    if "c3" in which:
        c3()
    torch.cuda.synchronize()
    print("ok")
