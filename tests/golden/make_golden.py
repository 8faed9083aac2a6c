"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (the reference lives at /root/reference and does not travel to
the GPU box):

    python tests/golden/make_golden.py

For every case it imports the reference's own functions --
    models.layers.warp                                   (scripts/models/layers.py:73)
    FlowInterpolationModel.compute_inputs                (scripts/models/flow_interpolation.py:338)
    FlowInterpolationModel.compute_output_image          (scripts/models/flow_interpolation.py:394)
-- runs forward and backward (autograd) on seeded inputs and stores inputs, outputs and
gradients.  It also checks, before writing, that oracle/torch_oracle.py reproduces the reference
bit for bit on CPU.  The fixtures are the pin for oracle/ssm_oracle.c (tests/test_oracle_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/scripts"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from models import layers as ref_layers                      # noqa: E402  (the reference)
from models.flow_interpolation import FlowInterpolationModel  # noqa: E402

from oracle import torch_oracle                               # noqa: E402
from ssm_b200 import synthetic                                # noqa: E402

torch.set_num_threads(8)


class _Ref:
    """Borrow the reference methods without building its U-Net (they do not touch self beyond
    `verbose` and `extract_outputs`)."""
    verbose = False
    compute_inputs = FlowInterpolationModel.compute_inputs
    extract_outputs = FlowInterpolationModel.extract_outputs
    compute_output_image = FlowInterpolationModel.compute_output_image


REF_MODEL = _Ref()

# name, B, H, W, flow kind, frame smoothness, t values
CASES = [
    ("small_smooth", 2, 24, 40, "smooth", True, [0.25, 0.625]),
    ("odd_size_noise", 1, 19, 37, "noise", False, [0.5]),
    ("zero_flow", 1, 16, 32, "zero", True, [0.125]),
    ("integer_flow", 1, 16, 32, "integer", False, [0.875]),
    ("border", 2, 20, 28, "border", True, [0.375, 0.75]),
    ("wide_1920", 1, 2, 1920, "smooth", False, [0.5]),
    ("tall_1088", 1, 1088, 2, "smooth", False, [0.625]),
]


def run_case(name, B, H, W, kind, smooth, tvals, seed):
    img6 = synthetic.frames(B, H, W, seed=seed, smooth=smooth)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=6.0, seed=seed + 1, kind=kind)
    out5 = synthetic.unet_out5(B, 1, H, W, seed=seed + 2)[:, 0].contiguous()
    t = torch.tensor(tvals, dtype=torch.float32).view(B, 1, 1, 1)
    g = torch.Generator().manual_seed(seed + 7)
    rec = {"img6": img6, "flow4": flow4, "out5": out5, "t": t.view(-1)}

    # ---- warp (layers.py:73) with autograd backward
    x = img6[:, 0:3].clone().requires_grad_(True)
    f = flow4[:, 0:2].clone().requires_grad_(True)
    y = ref_layers.warp(x, f)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    rec.update(warp_out=y.detach(), warp_gout=gy, warp_gimg=x.grad, warp_gflow=f.grad)
    x2 = img6[:, 0:3].clone().requires_grad_(True)
    f2 = flow4[:, 0:2].clone().requires_grad_(True)
    y2 = torch_oracle.warp(x2, f2)
    y2.backward(gy)
    assert torch.equal(y2, y) and torch.equal(x2.grad, x.grad) and torch.equal(f2.grad, f.grad), name

    # ---- compute_inputs (flow_interpolation.py:338)
    a = img6.clone().requires_grad_(True)
    b = flow4.clone().requires_grad_(True)
    in16 = REF_MODEL.compute_inputs(a, b, t)
    g16 = torch.randn(in16.shape, generator=g)
    in16.backward(g16)
    rec.update(in16=in16.detach(), pack_g16=g16, pack_gimg=a.grad, pack_gflow=b.grad)
    a2 = img6.clone().requires_grad_(True)
    b2 = flow4.clone().requires_grad_(True)
    o2 = torch_oracle.compute_inputs(a2, b2, t)
    o2.backward(g16)
    assert torch.equal(o2, in16) and torch.equal(a2.grad, a.grad) and torch.equal(b2.grad, b.grad), name

    # ---- compute_output_image (flow_interpolation.py:394)
    a = img6.clone().requires_grad_(True)
    xin = in16.detach().clone().requires_grad_(True)
    yo = out5.clone().requires_grad_(True)
    frame = REF_MODEL.compute_output_image(a, xin, yo, t)
    g3 = torch.randn(frame.shape, generator=g)
    frame.backward(g3)
    # the gradient of the 16-channel input is zero outside the flow channels 6:10: store those only
    assert xin.grad[:, :6].abs().max() == 0 and xin.grad[:, 10:].abs().max() == 0, name
    rec.update(frame=frame.detach(), fuse_g3=g3, fuse_gimg=a.grad, fuse_gflows=xin.grad[:, 6:10].contiguous(),
               fuse_gout5=yo.grad)
    a2 = img6.clone().requires_grad_(True)
    x2 = in16.detach().clone().requires_grad_(True)
    y2 = out5.clone().requires_grad_(True)
    f2 = torch_oracle.compute_output_image(a2, x2, y2, t)
    f2.backward(g3)
    assert torch.equal(f2, frame) and torch.equal(a2.grad, a.grad), name
    assert torch.equal(x2.grad, xin.grad) and torch.equal(y2.grad, yo.grad), name

    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.detach().numpy() for k, v in rec.items()})
    return rec


if __name__ == "__main__":
    total = 0
    for i, (name, B, H, W, kind, smooth, tvals) in enumerate(CASES):
        run_case(name, B, H, W, kind, smooth, tvals, seed=1000 + 17 * i)
        sz = os.path.getsize(os.path.join(HERE, name + ".npz"))
        total += sz
        print("%-16s B=%d %4dx%-4d %-8s -> %.1f KB" % (name, B, H, W, kind, sz / 1024))
    print("torch_oracle == reference bit for bit on every case; total %.2f MB" % (total / 2 ** 20))


# ---- the model loop (rows a6-a8): the reference's FullModel run on CPU ---------------------------
BOTTLENECK_CODE = {"CONV": 0, "CLSTM": 1, "CGRU": 2}


def run_loop_case(name, n_frames, seed, ini="superslomo_original.ini", bottleneck=None):
    """Runs scripts/models/superslomo_r.py::FullModel (bottleneck from the .ini -- CONV for
    superslomo_original.ini, CLSTM for superslomo_recurrent.ini -- or overridden; random-init U-Nets from a
    fixed seed, LAMBDA_P = 0 because the VGG weights cannot be downloaded) in inference and training
    mode on CPU.  The reference hard-codes .cuda() (superslomo_r.py:211); Tensor.cuda is patched to
    the identity for this run only.  U-Net weights are NOT stored (155 MB): tests rebuild them from
    the same seed (same construction order, same state_dict keys)."""
    import configparser
    import torchvision
    from models import superslomo_r as ref_model

    cfg = configparser.RawConfigParser()
    cfg.read("/root/reference/configs/" + ini)
    for sec in ("STAGE1", "STAGE2"):
        cfg.set(sec, "LOADPREV", "FALSE")
        cfg.set(sec, "FREEZE", "FALSE")
        if bottleneck is not None:
            cfg.set(sec, "BOTTLENECK", bottleneck)
    cfg.set("TRAIN", "LAMBDA_P", "0")
    cfg.set("TRAIN", "N_FRAMES", str(n_frames))
    real_vgg, real_cuda = torchvision.models.vgg16, torch.Tensor.cuda
    torchvision.models.vgg16 = lambda pretrained=True: real_vgg(weights=None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        torch.manual_seed(seed)
        model = ref_model.FullModel(cfg)
        B, H, W = 2, 64, 64
        frames = synthetic.frames(B, H, W, n_frames=n_frames, seed=seed + 1).view(B, n_frames, 3, H, W)
        targets = synthetic.frames(B, H, W, n_frames=n_frames - 1, seed=seed + 2).view(B, n_frames - 1, 3, H, W)
        t = synthetic.random_timesteps(B, n_frames - 1, seed=seed + 3).view(B, n_frames - 1, 1, 1, 1)
        with torch.no_grad():
            est, extras = model(frames, t, inference_mode=True)
        est_tr, losses = model(frames, t, target_images=targets, iteration=2, inference_mode=False)
        losses.mean(dim=0)[0].backward()
        g1 = model.stage1_model.final_conv.weight.grad.clone()
        g2 = model.stage2_model.final_conv.weight.grad.clone()
    finally:
        torchvision.models.vgg16, torch.Tensor.cuda = real_vgg, real_cuda
    rec = {"frames": frames, "targets": targets, "t": t.reshape(B, n_frames - 1), "est": est,
           "losses": losses.detach(), "est_train": est_tr.detach(), "grad_stage1_final": g1, "grad_stage2_final": g2,
           "seed": torch.tensor(seed), "bottleneck": torch.tensor(BOTTLENECK_CODE[cfg.get("STAGE1", "BOTTLENECK")])}
    for i, e in enumerate(extras):
        rec["extra%d" % i] = e
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.detach().numpy() for k, v in rec.items()})
    print("%-16s N_FRAMES=%d -> %.1f KB" % (name, n_frames, os.path.getsize(os.path.join(HERE, name + ".npz")) / 1024))


if __name__ == "__main__":
    run_loop_case("loop_frames2", 2, seed=4242)
    run_loop_case("loop_frames4", 4, seed=4343)
    # the recurrent configuration (C4): bidirectional ConvLSTM bottleneck over the 3 windows of 4 frames, and the
    # ConvGRU alternative the reference also builds (flow_computation.py:81-88)
    run_loop_case("loop_ssmr4_clstm", 4, seed=4444, ini="superslomo_recurrent.ini")
    run_loop_case("loop_ssmr3_cgru", 3, seed=4545, ini="superslomo_recurrent.ini", bottleneck="CGRU")


# ---- checkpoint layout: state_dict keys and shapes of the reference U-Nets, per bottleneck -------------------
def dump_state_dict_layout():
    """tests/golden/state_dict_layout.json: {bottleneck: {stage: [[key, shape], ...]}} of the reference's
    FlowComputationModel / FlowInterpolationModel (scripts/models/unetflow.py:11-32), so that the in-tree
    U-Nets can be checked to load the author's checkpoints key for key."""
    import configparser
    import json
    from models import unetflow

    out = {}
    for bottleneck in ("CONV", "CLSTM", "CGRU"):
        cfg = configparser.RawConfigParser()
        cfg.read("/root/reference/configs/superslomo_recurrent.ini")
        for sec in ("STAGE1", "STAGE2"):
            cfg.set(sec, "BOTTLENECK", bottleneck)
        out[bottleneck] = {}
        for stage, (cin, cout) in ((1, (6, 4)), (2, (16, 5))):
            model = unetflow.get_model(None, cin, cout, True, stage=stage, cfg=cfg)
            out[bottleneck]["stage%d" % stage] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    with open(os.path.join(HERE, "state_dict_layout.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("state_dict_layout.json:", {b: {s: len(v) for s, v in d.items()} for b, d in out.items()})
