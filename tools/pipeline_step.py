#!/usr/bin/env python
"""End-to-end inference step of the whole model on synthetic 1080p clips -- uint8 frames in pinned host
memory in, uint8 interpolated frames in pinned host memory out -- with the two U-Nets included
("report hot-path-only and end-to-end (with U-Nets) separately").  NOT the bench.py
headline: the U-Nets are stock torch/cuDNN modules with random-init weights (out of scope), and they
dominate this number.

    python tools/pipeline_step.py [--pairs 2] [--timesteps 7] [--height 1080] [--width 1920]
                                  [--amp] [--channels-last] [--graph] [--unet-chunk 1]

One step = H2D of the uint8 frames, ssm_frames_from_u8 (normalise + pad + RGBx), stage-1 U-Net once per
pair, compute_inputs for all N timesteps (one launch), stage-2 U-Net in chunks of --unet-chunk timesteps,
compute_output_image for all N timesteps (one launch), ssm_frames_to_u8 (crop + de-normalise), D2H.
With --graph the device part is captured once into a CUDA graph and replayed: the
C ABI never allocates or synchronises, so every call of the path is capturable.
Prints one JSON line: frames/s end to end, ms/step, and the share of the step spent in this repo's
kernels (events around every C-ABI call in a separate eager pass).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200.superslomo_r import FullModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--timesteps", type=int, default=7)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--unet-chunk", type=int, default=1, help="timesteps per stage-2 call (activation memory)")
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the U-Nets (the path stays fp32)")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--graph", action="store_true", help="capture the device part of the step into a CUDA graph")
    ap.add_argument("--generic-plumbing", action="store_true",
                    help="planar fp32 tensors either side of the stage-2 U-Net even when it runs channels-last / autocast "
                         "(default: compute_inputs writes the channels-last [bf16] tensor conv1a consumes)")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="let cuDNN time its algorithms per shape")
    ap.add_argument("--fp32-frames", action="store_true",
                    help="round-1 form: normalise to fp32 planes, FullModel.interpolate (RGBx gathers), frames_to_u8")
    ap.add_argument("--profile", default=None, help="write the kernel table of one step (torch.profiler) to this file")
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = a.cudnn_benchmark
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N, h_in, w_in = a.pairs, a.timesteps, a.height, a.width

    torch.manual_seed(42)
    model = FullModel(cfg=None, loss=ssm_b200.losses.SSMLosses(perceptual_features="zero")).to(dev).eval()   # inference only
    if a.channels_last:
        model.stage1_model.set_channels_last()
        model.stage2_model.set_channels_last()
    model.unet_layouts = not a.generic_plumbing
    # device-resident t: a host list would be copied from pageable memory inside the captured region
    t_values = torch.tensor([(k + 1) / (N + 1) for k in range(N)], dtype=torch.float32, device=dev)

    g = torch.Generator().manual_seed(1)
    h_u8 = torch.randint(0, 256, (B * 2, h_in, w_in, 3), dtype=torch.uint8, generator=g).pin_memory()
    d_u8 = torch.empty_like(h_u8, device=dev)
    h_out = torch.empty((B * N, h_in, w_in, 3), dtype=torch.uint8).pin_memory()
    lut = ssm_b200.normalisation_lut(device=dev)
    pad_values = lut[:, 0].tolist()           # read back once, outside the step

    def device_part():
        if not a.fp32_frames:       # 8-bit images straight through: byte-table warps, uint8 frames written by the fusion kernel
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
                out = model.interpolate_u8(d_u8.view(B, 2, h_in, w_in, 3), t_values, order="bgr", unet_chunk=a.unet_chunk)
            return out.view(B * N, h_in, w_in, 3)
        planar, _, (top, left) = ssm_b200.frames_from_u8(d_u8, order="bgr", pad_mode="before", lut=lut,
                                                           pad_values=pad_values)
        H, W = planar.shape[-2:]
        clip = planar.view(B, 2, 3, H, W)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            frames = model.interpolate(clip, t_values, unet_chunk=a.unet_chunk)       # B x N x 3 x H x W
        return ssm_b200.frames_to_u8(frames.reshape(B * N, 3, H, W), top=top, left=left, h_out=h_in, w_out=w_in,
                                     order="bgr", saturate=True)

    graph = None
    static_out = None

    def step():
        nonlocal static_out
        d_u8.copy_(h_u8, non_blocking=True)
        with torch.no_grad():
            if graph is not None:
                graph.replay()
                out = static_out
            else:
                out = device_part()
        h_out.copy_(out, non_blocking=True)

    with torch.no_grad():
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        if a.graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                device_part()                       # warm-up on the capture stream (cuDNN plans, allocator)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = device_part()
            step()
            torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    checksum = int(h_out[:, ::97, ::89].to(torch.int64).sum())

    # share of the step in this repo's kernels: events around every C-ABI call, eager pass
    spans, lib = [], ssm_b200._abi.lib()
    names = ["ssm_frames_from_u8", "ssm_frames_to_u8", "ssm_pack_frames", "ssm_flow_pack_fwd", "ssm_fuse_flow_fwd",
             "ssm_flow_pack_fwd_nhwc", "ssm_fuse_flow_fwd_mixed", "ssm_quads_from_u8", "ssm_flow_pack_fwd_q8",
             "ssm_flow_pack_fwd_q8_nhwc", "ssm_fuse_flow_fwd_q8", "ssm_fuse_flow_fwd_q8_u8"]
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

    class _Shim:
        def __getattr__(self, n):
            return timed(n, originals[n]) if n in originals else getattr(lib, n)

    ssm_b200._abi._lib = _Shim()
    with torch.no_grad():
        device_part()
    torch.cuda.synchronize()
    ssm_b200._abi._lib = lib
    path_ms = {}
    for n, s, e in spans:
        path_ms[n] = path_ms.get(n, 0.0) + s.elapsed_time(e)

    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof, torch.no_grad():
            device_part()
            torch.cuda.synchronize()
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))

    print(json.dumps({
        "what": "end-to-end inference step with the U-Nets (uint8 host frames -> uint8 host frames), not the bench headline",
        "pairs": B, "timesteps": N, "height": h_in, "width": w_in, "unet_chunk": a.unet_chunk,
        "amp_bf16_unets": a.amp, "channels_last_unets": a.channels_last, "cuda_graph": a.graph, "cudnn_benchmark": a.cudnn_benchmark,
        "unet_layouts": bool(model.unet_layouts and a.channels_last),
        "ms_per_step": ms, "frames_per_s": B * N / (ms * 1e-3), "checksum": checksum,
        "h2d_bytes_per_step": h_u8.numel(), "d2h_bytes_per_step": h_out.numel(),
        "path_kernels_ms": path_ms, "path_ms_total": sum(path_ms.values()),
        "path_share_of_step": sum(path_ms.values()) / ms,
        "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
