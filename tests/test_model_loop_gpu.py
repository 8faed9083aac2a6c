"""GPU suite: the timestep-batched model loop (rows a6-a8, a10) against
  * the reference's own FullModel outputs (golden, CPU run; U-Nets rebuilt from the seed), and
  * the window-by-window / timestep-by-timestep restatement of the reference loop run on the same
    device with the same U-Net modules (oracle/torch_oracle.py).
The U-Nets are stock torch modules (cuDNN convs, TF32 off for these comparisons); only the synthesis
path between and after them is ours.
"""
import copy

import pytest
import torch

import ssm_b200
from oracle import torch_oracle
from ssm_b200 import synthetic
from ssm_b200.superslomo_r import FullModel
from util import assert_close_fp32, load_golden, loop_cases, max_err, seeded_unets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _exact_convs():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.enabled)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.enabled = prev
    ssm_b200.set_coord_mode("cpu")


def _model(seed, bottleneck="CONV"):
    s1, s2 = seeded_unets(seed, DEV, bottleneck=bottleneck)
    return FullModel(cfg=None, stage1_model=s1, stage2_model=s2,
                     loss=ssm_b200.losses.SSMLosses(lambda_r=60.0, lambda_p=0.0, lambda_w=10.0))


@pytest.mark.parametrize("name", loop_cases())
def test_fullmodel_vs_reference_golden(name):
    """Reference FullModel (CPU) vs FullModel here (CUDA, coord mode "cpu").  The U-Net convolutions
    run in different libraries (oneDNN vs cuDNN), so flows differ by conv rounding (~1e-6) and the
    comparison is at 1e-4, not at the hot path's 1e-5."""
    d = load_golden(name)
    ssm_b200.set_coord_mode("cpu")
    m = _model(d["seed"].item(), d.get("bottleneck", 0))
    frames, t = d["frames"].to(DEV), d["t"].to(DEV)
    t5 = t.view(*t.shape, 1, 1, 1)
    with torch.no_grad():
        est, extras = m(frames, t5, inference_mode=True)
    assert_close_fp32(est, d["est"], "inference frame", tol=1e-4)
    assert len(extras) == 7
    for i, e in enumerate(extras):
        assert_close_fp32(e, d["extra%d" % i], "inference extra %d" % i, tol=1e-4)
    est_tr, losses = m(frames, t5, target_images=d["targets"].to(DEV), iteration=2, inference_mode=False)
    assert losses.shape == (frames.shape[0], 4)
    assert_close_fp32(est_tr, d["est_train"], "training frame", tol=1e-4)
    assert_close_fp32(losses, d["losses"], "losses [B,4]", tol=2e-3)
    losses.mean(dim=0)[0].backward()
    g1, g2 = m.stage1_model.final_conv.weight.grad, m.stage2_model.final_conv.weight.grad
    assert max_err(g1, d["grad_stage1_final"]) <= 2e-3 * max(1.0, d["grad_stage1_final"].abs().max().item())
    assert max_err(g2, d["grad_stage2_final"]) <= 2e-3 * max(1.0, d["grad_stage2_final"].abs().max().item())


@pytest.mark.parametrize("n_frames", [2, 4])
def test_fullmodel_vs_reference_loop_on_device(n_frames):
    """Same U-Net modules, same device: the reference loop restated with the reference's torch ops
    (ATen CUDA kernels: cuDNN off for grid_sample) vs the batched loop in coord mode "cuda"."""
    B, H, W = 2, 96, 128
    ssm_b200.set_coord_mode("cuda")
    m = _model(77 + n_frames)
    ref1, ref2 = copy.deepcopy(m.stage1_model), copy.deepcopy(m.stage2_model)
    frames = synthetic.frames(B, H, W, n_frames=n_frames, seed=5).view(B, n_frames, 3, H, W).to(DEV)
    targets = synthetic.frames(B, H, W, n_frames=n_frames - 1, seed=6).view(B, n_frames - 1, 3, H, W).to(DEV)
    t = synthetic.random_timesteps(B, n_frames - 1, seed=7).to(DEV).view(B, n_frames - 1, 1, 1, 1)
    torch.backends.cudnn.enabled = False      # grid_sample -> ATen kernel; convs -> ATen too, in both paths
    with torch.no_grad():
        est, extras = m(frames, t, inference_mode=True)
        r_est, r_extras = torch_oracle.model_forward(ref1, ref2, frames, t)
    assert_close_fp32(est, r_est, "inference frame")
    for i, (a, b) in enumerate(zip(extras, r_extras)):
        assert_close_fp32(a, b, "inference extra %d" % i)
    est_tr, losses = m(frames, t, target_images=targets, inference_mode=False)
    r_tr, r_losses = torch_oracle.model_forward(ref1, ref2, frames, t, target_images=targets)
    assert_close_fp32(est_tr, r_tr, "training frame")
    assert_close_fp32(losses, r_losses, "losses", tol=2e-4)
    losses.mean(dim=0)[0].backward()
    r_losses.mean(dim=0)[0].backward()
    for (n1, p1), (_, p2) in zip(m.stage1_model.named_parameters(), ref1.named_parameters()):
        scale = max(1.0, p2.grad.abs().max().item())
        assert max_err(p1.grad, p2.grad) <= 1e-3 * scale, "stage-1 grad %s" % n1
    for (n1, p1), (_, p2) in zip(m.stage2_model.named_parameters(), ref2.named_parameters()):
        scale = max(1.0, p2.grad.abs().max().item())
        assert max_err(p1.grad, p2.grad) <= 1e-3 * scale, "stage-2 grad %s" % n1


@pytest.mark.parametrize("n_frames,n_t,chunk,bottleneck", [(2, 7, None, "CONV"), (2, 7, 3, "CONV"), (4, 3, None, "CONV"),
                                                           (4, 7, None, "CLSTM"), (4, 3, 2, "CGRU")])
def test_interpolate_vs_per_timestep_reference_loop(n_frames, n_t, chunk, bottleneck):
    """interpolate(): stage 1 once, all N times per launch, vs the reference's loop that calls the
    whole model once per intermediate time (evaluate_interpolation_results.py:234-242).  CLSTM / CGRU: the
    recurrent configuration (C4, superslomo_recurrent.ini) -- the bottleneck couples the 3 windows of a sample, the
    N times are folded into the batch next to them."""
    B, H, W = 2, 64, 96
    ssm_b200.set_coord_mode("cuda")
    m = _model(99, bottleneck)
    frames = synthetic.frames(B, H, W, n_frames=n_frames, seed=15).view(B, n_frames, 3, H, W).to(DEV)
    torch.backends.cudnn.enabled = False
    tv = torch.arange(1, n_t + 1, dtype=torch.float32) / (n_t + 1)
    out = m.interpolate(frames, tv, unet_chunk=chunk)
    with torch.no_grad():
        ref = torch_oracle.interpolate_frames(m.stage1_model, m.stage2_model, frames, n_t)
    assert out.shape == (B, n_t, 3, H, W)
    # stage 2 runs on a differently shaped batch (B*N instead of B): allow conv rounding
    assert_close_fp32(out, ref, "interpolate vs per-timestep loop", tol=1e-4)


@pytest.mark.parametrize("n_frames,n_t,chunk,bottleneck,h,w", [(2, 7, None, "CONV", 60, 90), (2, 3, 2, "CONV", 64, 96),
                                                               (4, 3, None, "CLSTM", 45, 77)])
def test_interpolate_u8_matches_interpolate_on_normalised_frames(n_frames, n_t, chunk, bottleneck, h, w):
    """interpolate_u8 (8-bit images in, the warps gather raw bytes, uint8 images out) against interpolate() on the same
    images normalised and padded by ssm_frames_from_u8 -- i.e. against the path already pinned to the reference loop
    above.  Normalised frames within 1e-4 (the stage-2 U-Net sees inputs that differ by <= 3e-6); uint8 images equal
    to frames_to_u8 of the fp32 result up to one count where that result sits on a rounding boundary."""
    B = 2
    m = _model(123, bottleneck)
    torch.backends.cudnn.enabled = False
    g = torch.Generator().manual_seed(5)
    images = torch.randint(0, 256, (B, n_frames, h, w, 3), dtype=torch.uint8, generator=g).to(DEV)
    tv = torch.arange(1, n_t + 1, dtype=torch.float32) / (n_t + 1)
    planar, _, (top, left) = ssm_b200.frames_from_u8(images.view(B * n_frames, h, w, 3), order="bgr")
    H, W = planar.shape[-2:]
    ref = m.interpolate(planar.view(B, n_frames, 3, H, W), tv, unet_chunk=chunk)
    got = m.interpolate_u8(images, tv, order="bgr", unet_chunk=chunk, as_u8=False)
    assert got.shape == (B, n_t, 3, H, W)
    assert_close_fp32(got, ref, "interpolate_u8 (normalised frames)", tol=1e-4)
    got8 = m.interpolate_u8(images, tv, order="bgr", unet_chunk=chunk)
    ref8 = ssm_b200.frames_to_u8(ref.view(B * n_t, 3, H, W), top, left, h, w, order="bgr", saturate=True).view(B, n_t, h, w, 3)
    assert got8.shape == (B, n_t, h, w, 3) and got8.dtype == torch.uint8
    diff = (got8.int() - ref8.int()).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3
