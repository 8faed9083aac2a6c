"""Generates tests/golden/large_352_flow100.npz and tests/golden/sliding_window.json by running the UNMODIFIED
reference on CPU (build container only; /root/reference does not travel):

    python tests/golden/make_golden_large.py

large_352_flow100: one 352 x 352 pair (the training crop of BASELINE configs[2]) with flows of up to +-100 px
(control grid at 1/8 resolution x 20 px) through the reference's compute_inputs and
compute_output_image, forward and autograd backward.  To keep the file small the inputs are stored as integers
that reproduce the fp32 tensors exactly on any machine --
    frames  uint8 images, normalised with the reference's expression (visualize_interpolation.py:257-262)
    flow4   int16 / 64        (1/64 px steps)
    out5    int16 / 1024
    upstream gradients  int16 / 256, non-zero on a band of 96 rows
-- and the outputs (fused frame, the two warped images of compute_inputs, the flow / U-Net-output gradients of
compute_output_image, the flow gradient of compute_inputs) are stored for that band of rows only; the whole
352 x 352 problem is computed, and the band's samples reach far outside it.

sliding_window: the reference's own Interpolator.sliding_window (visualize_interpolation.py:270-288) for a range of
(n_images, n_frames, stride) -- the rule formats.sliding_window must follow.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/scripts"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from models.flow_interpolation import FlowInterpolationModel  # noqa: E402  (the reference)

from oracle import torch_oracle                               # noqa: E402
from ssm_b200 import synthetic                                # noqa: E402

torch.set_num_threads(8)

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # configs/superslomo_original.ini:57-58


class _Ref:
    verbose = False
    compute_inputs = FlowInterpolationModel.compute_inputs
    extract_outputs = FlowInterpolationModel.extract_outputs
    compute_output_image = FlowInterpolationModel.compute_output_image


def inputs_from_integers(u8, flow_i16, out5_i16):
    """the exact fp32 inputs of the fixture from its stored integers (also used by the tests)"""
    x = u8.permute(0, 3, 1, 2).float() / 255.0
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    x = (x - mean) / std                                      # visualize_interpolation.py:257-262
    img6 = torch.cat([x[0::2], x[1::2]], dim=1).contiguous()  # pair b = images 2b, 2b+1
    return img6, flow_i16.float() / 64.0, out5_i16.float() / 1024.0


BAND = (128, 224)       # rows whose outputs are stored (the whole 352 x 352 problem is computed)


def upstream_from_integers(g16_i16, g3_i16):
    """full-size upstream gradients of compute_inputs (16 channels; 3:13 stored) and compute_output_image"""
    g16 = torch.zeros((g16_i16.shape[0], 16) + tuple(g16_i16.shape[2:]))
    g16[:, 3:13] = g16_i16.float() / 256.0
    return g16, g3_i16.float() / 256.0


def make_large(name="large_352_flow100", H=352, W=352, seed=7100):
    B = 1
    g = torch.Generator().manual_seed(seed)
    smooth = synthetic.frames(2 * B, H, W, n_frames=1, seed=seed, smooth=True)
    smooth = (smooth - smooth.amin()) / (smooth.amax() - smooth.amin())
    u8 = (smooth.permute(0, 2, 3, 1) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous()
    flow = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=seed + 1, kind="smooth")
    flow_i16 = (flow * 64.0).round().clamp(-32768, 32767).to(torch.int16)
    out5 = synthetic.unet_out5(B, 1, H, W, seed=seed + 2)[:, 0]
    out5_i16 = (out5 * 1024.0).round().clamp(-32768, 32767).to(torch.int16)
    img6, flow4, y5 = inputs_from_integers(u8, flow_i16, out5_i16)
    t = torch.tensor([0.375]).view(B, 1, 1, 1)
    ref = _Ref()

    # upstream gradients: multiples of 1/256 (int16), non-zero on the stored band of rows only -- the flow and
    # U-Net-output gradients at a pixel depend on the upstream gradient at that pixel alone
    r0, r1 = BAND
    g16_i16 = torch.zeros((B, 10, H, W), dtype=torch.int16)
    g3_i16 = torch.zeros((B, 3, H, W), dtype=torch.int16)
    g16_i16[:, :, r0:r1] = (torch.randn((B, 10, r1 - r0, W), generator=g) * 256.0).round().to(torch.int16)
    g3_i16[:, :, r0:r1] = (torch.randn((B, 3, r1 - r0, W), generator=g) * 256.0).round().to(torch.int16)
    g16, g3 = upstream_from_integers(g16_i16, g3_i16)

    a, b = img6.clone(), flow4.clone().requires_grad_(True)
    in16 = ref.compute_inputs(a, b, t)
    in16.backward(g16)
    assert torch.equal(torch_oracle.compute_inputs(img6, flow4, t), in16)

    xin, yo = in16.detach().clone().requires_grad_(True), y5.clone().requires_grad_(True)
    frame = ref.compute_output_image(img6, xin, yo, t)
    frame.backward(g3)
    assert torch.equal(torch_oracle.compute_output_image(img6, in16.detach(), y5, t), frame)

    def band(x):
        return x.detach()[:, :, r0:r1].contiguous()
    rec = {"u8": u8, "flow_i16": flow_i16, "out5_i16": out5_i16, "t": t.view(-1), "band": torch.tensor(BAND),
           "flow_absmax": flow4.abs().max().view(1),
           "g16_band_i16": g16_i16[:, :, r0:r1].contiguous(), "g3_band_i16": g3_i16[:, :, r0:r1].contiguous(),
           "in16_warped": band(torch.cat([in16[:, 3:6], in16[:, 10:13]], 1)), "pack_gflow": band(b.grad),
           "frame": band(frame), "fuse_gflows": band(xin.grad[:, 6:10]), "fuse_gout5": band(yo.grad)}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in rec.items()})
    print("%s: %dx%d, max |flow| %.1f px, outputs of rows %d:%d -> %.2f MB" % (
        name, H, W, flow4.abs().max().item(), r0, r1, os.path.getsize(path) / 2 ** 20))


def make_sliding_window():
    """Interpolator.sliding_window reads self.n_frames and self.is_fps_240 only: call it unbound on a bare object."""
    for mod in ("tensorboardX", "skimage", "skimage.measure", "skimage.metrics", "more_itertools", "matplotlib", "matplotlib.pyplot"):
        if mod not in sys.modules:
            try:
                __import__(mod)
            except Exception:
                import types
                m = types.ModuleType(mod)
                m.__path__ = []
                sys.modules[mod] = m
    import importlib
    vis = importlib.import_module("visualize_interpolation")
    cls = vis.Interpolator
    out = []
    for n_images in (1, 2, 3, 4, 5, 9, 16, 33):
        for n_frames in (2, 4, 6):
            for fps240 in (False, True):
                class Bare:
                    pass
                obj = Bare()
                obj.n_frames = n_frames
                obj.is_fps_240 = fps240
                seq = list(range(n_images))      # "paths" = indices, so the yielded samples are the index lists
                windows = [list(w) for w in cls.sliding_window(obj, seq)]
                out.append({"n_images": n_images, "n_frames": n_frames, "is_fps_240": fps240, "windows": windows})
    path = os.path.join(HERE, "sliding_window.json")
    with open(path, "w") as f:
        json.dump({"source": "scripts/visualize_interpolation.py Interpolator.sliding_window", "cases": out}, f)
    print("sliding_window.json: %d cases" % len(out))


if __name__ == "__main__":
    make_large()
    make_sliding_window()
