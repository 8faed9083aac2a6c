"""Generates tests/golden/q8_frames.npz (and q8_border.npz): the UNMODIFIED reference on uint8 images.

Run in the build container only (the reference lives at /root/reference and does not travel):

    python tests/golden/make_golden_q8.py

The reference's real-video path reads uint8 images, pads them to a multiple of 32 with byte 0 and normalises them
(scripts/visualize_interpolation.py:61-88 load_batch, :257-262 normalize_tensor) before compute_inputs /
compute_output_image (scripts/models/flow_interpolation.py:338-429) see them.  This script runs exactly those
reference functions on seeded uint8 images written as PNG and read back by the reference's own cv2.imread, and
stores the images, flows, U-Net-output surrogate, times and the reference's results.  The fixtures pin the 8-bit
entry points (ssm_quads_from_u8, ssm_flow_pack_fwd_q8, ssm_fuse_flow_fwd_q8[_u8]), which gather raw bytes and
normalise after interpolating, against the reference's normalise-then-interpolate order.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden_frames as mgf          # noqa: E402  (registers the stand-ins for absent unrelated imports)

from models.flow_interpolation import FlowInterpolationModel  # noqa: E402  (the reference)

from ssm_b200 import synthetic            # noqa: E402


class _Ref:
    verbose = False
    compute_inputs = FlowInterpolationModel.compute_inputs
    extract_outputs = FlowInterpolationModel.extract_outputs
    compute_output_image = FlowInterpolationModel.compute_output_image


def run(name, B, h_in, w_in, kind, flow_px, tvals, seed, smooth_images):
    import cv2
    vis = mgf._import_reference()["visualize_interpolation"]
    rng = np.random.RandomState(seed)
    if smooth_images:       # low-passed noise, as a natural image is
        x = synthetic.frames(2 * B, h_in, w_in, n_frames=1, seed=seed, smooth=True)
        x = (x - x.amin()) / (x.amax() - x.amin())
        bgr = (x.permute(0, 2, 3, 1).numpy() * 255.0).round().astype(np.uint8)
    else:
        bgr = rng.randint(0, 256, size=(2 * B, h_in, w_in, 3)).astype(np.uint8)
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i in range(2 * B):
            p = os.path.join(d, "%05d.png" % i)
            assert cv2.imwrite(p, bgr[i])
            paths.append(p)
        bare = types.SimpleNamespace()
        loaded = vis.Interpolator.load_batch(bare, paths)                  # 1 x 2B x 3 x H x W, 0..255, padded
    normalised = vis.Interpolator.normalize_tensor(bare, loaded)[0]         # 2B x 3 x H x W
    H, W = normalised.shape[-2:]
    img6 = normalised.reshape(B, 6, H, W).contiguous()                      # pairs of consecutive frames
    flow4 = synthetic.flows(B, H, W, 4, flow_px=flow_px, seed=seed + 1, kind=kind)
    N = len(tvals)
    out5 = synthetic.unet_out5(B, N, H, W, seed=seed + 2)
    ref = _Ref()
    in16, frames = [], []
    for n, tv in enumerate(tvals):
        t = torch.full((B, 1, 1, 1), tv, dtype=torch.float32)
        x16 = ref.compute_inputs(img6, flow4, t)
        in16.append(x16)
        frames.append(ref.compute_output_image(img6, x16, out5[:, n], t))
    rec = {"bgr_u8": bgr, "img6": img6.numpy(), "flow4": flow4.numpy(), "out5": out5.numpy(),
           "t": np.asarray(tvals, dtype=np.float32), "in16": torch.stack(in16, 1).numpy(),
           "frames": torch.stack(frames, 1).numpy()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print("%-12s B=%d %dx%d -> %dx%d  %s  %.1f KB" % (name, B, h_in, w_in, H, W, kind, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    torch.set_num_threads(8)
    run("q8_frames", 2, 45, 70, "smooth", 6.0, [0.125, 0.5, 0.875], seed=7001, smooth_images=True)
    run("q8_border", 1, 40, 62, "border", 6.0, [0.375, 0.75], seed=7002, smooth_images=False)
