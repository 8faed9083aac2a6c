"""CPU suite: the oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py imported /root/reference/scripts and ran it on CPU)."""
import pytest
import torch

from oracle import c_oracle, torch_oracle
from util import assert_close_fp32, golden_cases, load_golden

CASES = golden_cases()


def test_golden_fixtures_present():
    assert len(CASES) >= 7


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference(name):
    d = load_golden(name)
    img3, flow2 = d["img6"][:, 0:3].contiguous(), d["flow4"][:, 0:2].contiguous()
    assert_close_fp32(c_oracle.warp(img3, flow2), d["warp_out"], "warp fwd")
    gi, gf = c_oracle.warp_backward(d["warp_gout"], img3, flow2)
    assert_close_fp32(gi, d["warp_gimg"], "warp grad img")
    assert_close_fp32(gf, d["warp_gflow"], "warp grad flow")

    in16 = c_oracle.compute_inputs(d["img6"], d["flow4"], d["t"])
    # the estimated flows fix the sampling coordinates downstream: they must be bit-identical
    assert torch.equal(in16[:, 6:10], d["in16"][:, 6:10]), "estimated flows are not bit-identical"
    assert torch.equal(in16[:, 0:3], d["in16"][:, 0:3]) and torch.equal(in16[:, 13:16], d["in16"][:, 13:16])
    assert_close_fp32(in16, d["in16"], "compute_inputs fwd")
    gi, gf = c_oracle.compute_inputs_backward(d["pack_g16"], d["img6"], d["flow4"], d["t"])
    assert_close_fp32(gi, d["pack_gimg"], "compute_inputs grad img")
    assert_close_fp32(gf, d["pack_gflow"], "compute_inputs grad flow")

    frame = c_oracle.compute_output_image(d["img6"], d["in16"], d["out5"], d["t"])
    assert_close_fp32(frame, d["frame"], "compute_output_image fwd")
    gi, gx, gy = c_oracle.compute_output_image_backward(d["fuse_g3"], d["img6"], d["in16"], d["out5"], d["t"])
    assert_close_fp32(gi, d["fuse_gimg"], "compute_output_image grad img")
    assert_close_fp32(gx[:, 6:10], d["fuse_gflows"], "compute_output_image grad in16[6:10]")
    assert gx[:, :6].abs().max() == 0 and gx[:, 10:].abs().max() == 0
    assert_close_fp32(gy, d["fuse_gout5"], "compute_output_image grad out5")


def test_c_oracle_matches_reference_large_flows():
    """352 x 352 with flows of up to 82 px (tests/golden/make_golden_large.py): the C restatement against the
    reference's own outputs and autograd gradients on the stored band of rows."""
    from util import load_large_golden
    d = load_large_golden()
    assert d["flow_absmax"].item() > 60.0
    rows = d["rows"]
    in16 = c_oracle.compute_inputs(d["img6"], d["flow4"], d["t"])
    warped = torch.cat([in16[:, 3:6], in16[:, 10:13]], 1)[:, :, rows]
    assert_close_fp32(warped, d["in16_warped"], "compute_inputs warped images, large flows")
    _, gf = c_oracle.compute_inputs_backward(d["g16"], d["img6"], d["flow4"], d["t"])
    assert_close_fp32(gf[:, :, rows], d["pack_gflow"], "compute_inputs grad flow, large flows")
    frame = c_oracle.compute_output_image(d["img6"], in16, d["out5"], d["t"])
    assert_close_fp32(frame[:, :, rows], d["frame"], "compute_output_image, large flows")
    _, gx, gy = c_oracle.compute_output_image_backward(d["g3"], d["img6"], in16, d["out5"], d["t"])
    assert_close_fp32(gx[:, 6:10, rows], d["fuse_gflows"], "compute_output_image grad in16[6:10], large flows")
    assert_close_fp32(gy[:, :, rows], d["fuse_gout5"], "compute_output_image grad out5, large flows")


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_matches_reference(name):
    """The torch restatement issues the reference's torch ops: same bits on the same torch build."""
    d = load_golden(name)
    t = d["t"].view(-1, 1, 1, 1)
    assert_close_fp32(torch_oracle.warp(d["img6"][:, 0:3], d["flow4"][:, 0:2]), d["warp_out"], "warp", tol=1e-6)
    assert_close_fp32(torch_oracle.compute_inputs(d["img6"], d["flow4"], t), d["in16"], "compute_inputs", tol=1e-6)
    assert_close_fp32(torch_oracle.compute_output_image(d["img6"], d["in16"], d["out5"], t), d["frame"],
                      "compute_output_image", tol=1e-6)


def test_coord_modes_differ_only_slightly():
    """RCP (CUDA-style reciprocal multiply) and DIV (CPU-style division) are two reference
    bit-patterns: close in the forward, not identical (SURVEY.md finding 3b)."""
    d = load_golden("wide_1920")
    img3, flow2 = d["img6"][:, 0:3].contiguous(), d["flow4"][:, 0:2].contiguous()
    a = c_oracle.warp(img3, flow2, coord_mode=c_oracle.COORD_DIV)
    b = c_oracle.warp(img3, flow2, coord_mode=c_oracle.COORD_RCP)
    assert (a - b).abs().max() < 5e-3
    assert not torch.equal(a, b)


def test_oracle_properties():
    """Analytic properties the reference satisfies (SURVEY.md section 8(c) iii)."""
    from ssm_b200 import synthetic
    B, H, W = 2, 32, 48
    img6 = synthetic.frames(B, H, W, seed=3)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=4.0, seed=4)
    out5 = synthetic.unet_out5(B, 1, H, W, seed=5)[:, 0].contiguous()
    t = torch.tensor([0.25, 0.625])
    in16 = c_oracle.compute_inputs(img6, flow4, t)
    # 16-channel layout (flow_interpolation.py:364-367)
    assert torch.equal(in16[:, 0:3], img6[:, 3:6]) and torch.equal(in16[:, 13:16], img6[:, 0:3])
    assert torch.equal(in16[:, 3:6], c_oracle.warp(img6[:, 3:6].contiguous(), in16[:, 6:8].contiguous()))
    assert torch.equal(in16[:, 10:13], c_oracle.warp(img6[:, 0:3].contiguous(), in16[:, 8:10].contiguous()))
    # time-reversal symmetry of the fusion
    frame = c_oracle.compute_output_image(img6, in16, out5, t)
    img_sw = torch.cat([img6[:, 3:6], img6[:, 0:3]], 1)
    in_sw = in16.clone()
    in_sw[:, 6:8], in_sw[:, 8:10] = in16[:, 8:10], in16[:, 6:8]
    out_sw = torch.cat([-out5[:, 0:1], out5[:, 3:5], out5[:, 1:3]], 1)
    frame_sw = c_oracle.compute_output_image(img_sw, in_sw, out_sw, 1.0 - t)
    assert (frame - frame_sw).abs().max() < 2e-6
    # saturated visibility stays finite
    sat = out5.clone()
    sat[:, 0] = 40.0
    assert torch.isfinite(c_oracle.compute_output_image(img6, in16, sat, t)).all()


# ---- model loop (rows a6-a8) ---------------------------------------------------------------------
from util import loop_cases, seeded_unets  # noqa: E402


@pytest.mark.parametrize("name", loop_cases())
def test_loop_oracle_matches_reference_fullmodel(name):
    """oracle.torch_oracle.model_forward (window-by-window restatement of superslomo_r.py:250-293)
    against the reference's own FullModel run on CPU (tests/golden/make_golden.py::run_loop_case).
    Also pins that the in-tree U-Nets rebuild the reference's weights from the seed."""
    d = load_golden(name)
    s1, s2 = seeded_unets(d["seed"].item(), bottleneck=d.get("bottleneck", 0))
    for m in (s1, s2):              # recurrent bottleneck: the reference's joint gate convolution, for its rounding
        if hasattr(m.conv6, "split_input"):
            m.conv6.split_input = False
    t = d["t"].view(*d["t"].shape, 1, 1, 1)
    with torch.no_grad():
        est, extras = torch_oracle.model_forward(s1, s2, d["frames"], t)
    assert_close_fp32(est, d["est"], "inference frame", tol=1e-6)
    for i, e in enumerate(extras):
        assert_close_fp32(e, d["extra%d" % i], "inference extra %d" % i, tol=1e-6)
    est_tr, losses = torch_oracle.model_forward(s1, s2, d["frames"], t, target_images=d["targets"],
                                                )
    assert_close_fp32(est_tr, d["est_train"], "training frame", tol=1e-6)
    assert_close_fp32(losses, d["losses"], "losses [B,4]", tol=2e-5)
    losses.mean(dim=0)[0].backward()
    assert_close_fp32(s1.final_conv.weight.grad, d["grad_stage1_final"], "stage-1 final_conv grad", tol=1e-5)
    assert_close_fp32(s2.final_conv.weight.grad, d["grad_stage2_final"], "stage-2 final_conv grad", tol=1e-5)


@pytest.mark.parametrize("name", [n for n in loop_cases() if "ssmr" in n])
def test_recurrent_bottleneck_split_gate_convolution(name):
    """The batched-input form of the ConvLSTM / ConvGRU bottleneck (recurrent.py: input half of every gate
    convolution run once over all windows) against the reference's FullModel: same weights, same result up to
    the summation order inside the convolutions."""
    d = load_golden(name)
    s1, s2 = seeded_unets(d["seed"].item(), bottleneck=d["bottleneck"])
    assert s1.conv6.split_input and s2.conv6.split_input
    t = d["t"].view(*d["t"].shape, 1, 1, 1)
    with torch.no_grad():
        est, extras = torch_oracle.model_forward(s1, s2, d["frames"], t)
    assert_close_fp32(est, d["est"], "inference frame", tol=1e-4)
    for i, e in enumerate(extras):
        assert_close_fp32(e, d["extra%d" % i], "inference extra %d" % i, tol=1e-4)


# ---- pre/post frame steps: restatements pinned to the reference's own methods --------
def _prepost():
    import os
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_prepost.npz"))


def test_prepost_oracle_matches_reference():
    """tests/golden/make_golden_frames.py ran Interpolator.load_batch/normalize_tensor, Evaluator's
    crop + denormalise + astype(uint8) and the data loader's Normalize/ToTensor/EvalPad unmodified on CPU;
    the restatements in oracle/torch_oracle.py must reproduce them bit for bit."""
    import numpy as np
    d = _prepost()
    got = torch_oracle.load_batch_and_normalize(d["vis_bgr_u8"])
    assert torch.equal(got, torch.from_numpy(d["vis_normalised"]))
    h0, w0, h, w = [int(v) for v in d["eval_crop"]]
    assert np.array_equal(torch_oracle.crop_denormalize_u8(torch.from_numpy(d["eval_in"]), h0, w0, h, w), d["eval_u8"])
    got = torch_oracle.reader_normalize_and_pad(d["reader_rgb_u8"], int(d["reader_pad"][0]))
    assert torch.equal(got, torch.from_numpy(d["reader_out"]))


# ---- the C restatement against the torch restatement (== the reference's op sequence) on random small cases ------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None)
@given(B=st.integers(1, 3), H=st.integers(1, 9), W=st.integers(1, 11), seed=st.integers(0, 2 ** 20),
       flow_scale=st.sampled_from([0.0, 0.3, 2.0, 40.0]), k=st.integers(1, 7))
def test_c_oracle_matches_torch_oracle_on_random_small_cases(B, H, W, seed, flow_scale, k):
    """Degenerate sizes (H or W = 1: the normalisation divides by max(size-1, 1)), flows that leave the image on every
    side, all timesteps k/8: forward values and the flow / logit / residual-flow gradients of the C oracle against
    autograd through the reference's torch ops."""
    g = torch.Generator().manual_seed(seed)
    img6 = torch.randn((B, 6, H, W), generator=g)
    flow4 = (torch.randn((B, 4, H, W), generator=g) * flow_scale).requires_grad_()
    out5 = torch.randn((B, 5, H, W), generator=g).requires_grad_()
    t = torch.full((B,), k / 8.0)
    t4 = t.view(B, 1, 1, 1)
    in16 = torch_oracle.compute_inputs(img6, flow4, t4)
    frame = torch_oracle.compute_output_image(img6, in16, out5, t4)
    c16 = c_oracle.compute_inputs(img6, flow4.detach(), t)
    assert torch.equal(c16[:, 6:10], in16[:, 6:10].detach())
    assert_close_fp32(c16, in16, "compute_inputs", tol=2e-6 * max(1.0, img6.abs().max().item()))
    cfr = c_oracle.compute_output_image(img6, c16, out5.detach(), t)
    # the final division amplifies rounding where the visibility-weighted denominator is small
    assert_close_fp32(cfr, frame, "compute_output_image", tol=1e-5 * max(1.0, img6.abs().max().item()))
    if flow_scale <= 2.0:       # gradients w.r.t. the flow are discontinuous at cell borders: compare where they are tame
        g3 = torch.randn(frame.shape, generator=g)
        gflow, gout = torch.autograd.grad(frame, (flow4, out5), g3)
        _, gx, gy = c_oracle.compute_output_image_backward(g3, img6, c16, out5.detach(), t, need_img=False)
        _, gf = c_oracle.compute_inputs_backward(gx, img6, flow4.detach(), t, need_img=False)
        scale = max(1.0, gflow.abs().max().item(), gout.abs().max().item())
        assert_close_fp32(gy, gout, "grad out5", tol=2e-5 * scale)
        assert_close_fp32(gf, gflow, "grad flow", tol=2e-5 * scale)


# ---- 8-bit frames: the reference's load_batch + normalize_tensor in front of the path ------------------------
from util import q8_cases  # noqa: E402


@pytest.mark.parametrize("name", q8_cases())
def test_oracle_on_uint8_frames_matches_reference(name):
    """tests/golden/make_golden_q8.py ran the reference's own image loading, padding, normalisation,
    compute_inputs and compute_output_image on uint8 images; the oracle restatements reproduce it."""
    d = load_golden(name)
    B = d["flow4"].shape[0]
    img = torch_oracle.load_batch_and_normalize(d["bgr_u8"].numpy())[0]               # 2B x 3 x H x W
    img6 = img.reshape(B, 6, *img.shape[-2:]).contiguous()
    assert torch.equal(img6, d["img6"]), "normalised frames are not bit-identical"
    for n, tv in enumerate(d["t"].tolist()):
        t = torch.full((B,), tv)
        in16 = c_oracle.compute_inputs(img6, d["flow4"], t)
        assert_close_fp32(in16, d["in16"][:, n], "compute_inputs n=%d" % n)
        frame = c_oracle.compute_output_image(img6, d["in16"][:, n].contiguous(), d["out5"][:, n].contiguous(), t)
        assert_close_fp32(frame, d["frames"][:, n], "compute_output_image n=%d" % n)
