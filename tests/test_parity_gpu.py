"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against
  * the golden vectors produced by the reference itself (tests/golden),
  * the C oracle (oracle/ssm_oracle.c) on seeded inputs, in both coordinate modes,
  * the reference's own torch ops run on the CUDA device (oracle/torch_oracle.py with cuDNN off =
    CUDA-ATen bit-pattern; with cuDNN on = cuDNN bit-pattern, reported, not asserted at 1e-5),
and through size-independent properties at BASELINE.json's full size (16 pairs 1088x1920, N=7).
Tolerances: fp32 max abs error <= 1e-5; bf16 storage <= 2e-2 (+ 2^-7 relative above magnitude 1).
"""
import pytest
import torch

import ssm_b200
from oracle import c_oracle, torch_oracle
from ssm_b200 import synthetic
from util import assert_close_bf16, assert_close_fp32, assert_close_scaled, assert_sum_close, golden_cases, load_golden, max_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODES = [("cpu", c_oracle.COORD_DIV), ("cuda", c_oracle.COORD_RCP)]


def _dev(x, grad=False):
    return x.to(DEV).requires_grad_(grad)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases())
def test_golden_vectors(name):
    """Reference outputs (CPU run of the unmodified reference) vs the CUDA path, mode "cpu"."""
    d = load_golden(name)
    B = d["img6"].shape[0]
    t = d["t"].view(B, 1, 1, 1)
    # warp fwd + bwd
    x, f = _dev(d["img6"][:, 0:3], True), _dev(d["flow4"][:, 0:2], True)
    y = ssm_b200.warp(x, f)
    y.backward(_dev(d["warp_gout"]))
    assert_close_fp32(y, d["warp_out"], "warp fwd")
    assert_close_fp32(x.grad, d["warp_gimg"], "warp grad img")
    assert_close_fp32(f.grad, d["warp_gflow"], "warp grad flow")
    # compute_inputs fwd + bwd through the reference-shaped method
    ops = ssm_b200.SynthesisMixin()
    a, b = _dev(d["img6"], True), _dev(d["flow4"], True)
    in16 = ops.compute_inputs(a, b, _dev(t))
    in16.backward(_dev(d["pack_g16"]))
    assert torch.equal(in16[:, 6:10].detach().cpu(), d["in16"][:, 6:10]), "estimated flows not bit-identical"
    assert_close_fp32(in16, d["in16"], "compute_inputs fwd")
    assert_close_fp32(a.grad, d["pack_gimg"], "compute_inputs grad img")
    assert_close_fp32(b.grad, d["pack_gflow"], "compute_inputs grad flow")
    # compute_output_image fwd + bwd
    a, xin, yo = _dev(d["img6"], True), _dev(d["in16"], True), _dev(d["out5"], True)
    frame = ops.compute_output_image(a, xin, yo, _dev(t))
    frame.backward(_dev(d["fuse_g3"]))
    assert_close_fp32(frame, d["frame"], "compute_output_image fwd")
    assert_close_fp32(a.grad, d["fuse_gimg"], "compute_output_image grad img")
    assert_close_fp32(xin.grad[:, 6:10], d["fuse_gflows"], "compute_output_image grad in16[6:10]")
    assert xin.grad[:, :6].abs().max() == 0 and xin.grad[:, 10:].abs().max() == 0
    assert_close_fp32(yo.grad, d["fuse_gout5"], "compute_output_image grad out5")


# ---------------------------------------------------------------------------------------------
def _inputs(B, N, H, W, seed, kind="smooth", smooth=True, flow_px=8.0):
    img6 = synthetic.frames(B, H, W, seed=seed, smooth=smooth)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=flow_px, seed=seed + 1, kind=kind)
    out5 = synthetic.unet_out5(B, N, H, W, seed=seed + 2)
    t = synthetic.timesteps(B, N) if N > 1 else synthetic.random_timesteps(B, 1, seed=seed + 3)
    return img6, flow4, out5, t


@pytest.mark.parametrize("mode_name,mode", MODES)
@pytest.mark.parametrize("B,N,H,W,kind,smooth", [
    (2, 3, 64, 96, "smooth", True),
    (1, 7, 45, 77, "noise", False),      # ragged: not a multiple of the 32x8 tile
    (2, 1, 33, 31, "border", True),      # narrower than one tile, samples cross every border
    (1, 2, 8, 1920, "integer", False),
    (1, 1, 1, 1, "zero", False),         # degenerate 1x1 frame: max(W-1,1) path
    (2, 2, 1, 40, "border", False),      # a single row: the y normalisation divides by max(H-1, 1) = 1
    (2, 2, 37, 1, "noise", False),       # a single column
    (3, 2, 352, 352, "smooth", True),    # training crop size
])
def test_batched_kernels_vs_c_oracle(mode_name, mode, B, N, H, W, kind, smooth):
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=100 + H + W, kind=kind, smooth=smooth)
    gen = torch.Generator().manual_seed(5)
    g16 = torch.randn(B, N, 16, H, W, generator=gen)
    g3 = torch.randn(B, N, 3, H, W, generator=gen)
    a, b = _dev(img6, True), _dev(flow4, True)
    in16 = ssm_b200.flow_pack(a, b, _dev(t), n_timesteps=N, coord_mode=mode_name)
    in16.backward(_dev(g16))
    a2, xin, yo = _dev(img6, True), _dev(in16.detach().cpu(), True), _dev(out5, True)
    frames = ssm_b200.fuse(a2, xin, yo, _dev(t), coord_mode=mode_name)
    frames.backward(_dev(g3))
    ref_gimg, ref_gflow, ref_gimg2 = [], [], []          # per-timestep terms of the sums the kernels accumulate
    for n in range(N):
        tn = t[:, n]
        r16 = c_oracle.compute_inputs(img6, flow4, tn, coord_mode=mode)
        assert torch.equal(in16[:, n, 6:10].detach().cpu(), r16[:, 6:10]), "estimated flows not bit-identical"
        assert_close_fp32(in16[:, n], r16, "flow_pack fwd n=%d" % n)
        gi, gf = c_oracle.compute_inputs_backward(g16[:, n].contiguous(), img6, flow4, tn, coord_mode=mode)
        ref_gimg.append(gi)
        ref_gflow.append(gf)
        y5 = out5[:, n].contiguous()
        r3 = c_oracle.compute_output_image(img6, r16, y5, tn, coord_mode=mode)
        assert_close_fp32(frames[:, n], r3, "fuse fwd n=%d" % n)
        gi2, gx, gy = c_oracle.compute_output_image_backward(g3[:, n].contiguous(), img6, r16, y5, tn, coord_mode=mode)
        ref_gimg2.append(gi2)
        assert_close_fp32(xin.grad[:, n], gx, "fuse grad in16 n=%d" % n)
        assert_close_fp32(yo.grad[:, n], gy, "fuse grad out5 n=%d" % n)
    # sums over the N timesteps: 1e-5 against the float64 sum of the oracle's per-timestep terms, plus the
    # representation bound of an fp32 running sum (util.assert_sum_close prints the reference's own summation noise)
    assert_sum_close(b.grad, ref_gflow, "flow_pack grad flow")
    assert_sum_close(a.grad, ref_gimg, "flow_pack grad img")
    assert_sum_close(a2.grad, ref_gimg2, "fuse grad img")


@pytest.mark.parametrize("mode_name,mode", MODES)
@pytest.mark.parametrize("B,N,H,W,kind,dtype", [
    (2, 3, 64, 96, "smooth", torch.float32),
    (1, 7, 45, 77, "border", torch.float32),
    (2, 1, 352, 352, "smooth", torch.float32),
    (2, 3, 64, 96, "smooth", torch.bfloat16),
])
def test_fuse_from_flow_equals_two_step_path(mode_name, mode, B, N, H, W, kind, dtype):
    """ssm_fuse_flow_fwd/bwd recompute input_tensor[:, 6:10] from flow_pred_tensor: the frames must be
    bit-identical to fuse(flow_pack(...)), the out5 gradient too, and the flow gradient must equal the
    chain through the 16-channel tensor (and the C oracle's) to rounding."""
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=300 + H + N, kind=kind)
    g3 = torch.randn(B, N, 3, H, W, generator=torch.Generator().manual_seed(8))
    cast = lambda x, grad=False: x.to(DEV).to(dtype).requires_grad_(grad)
    a, f, y, td = cast(img6, True), cast(flow4, True), cast(out5, True), _dev(t)
    frames = ssm_b200.fuse_from_flow(a, f, y, td, coord_mode=mode_name)
    frames.backward(cast(g3))
    a2, f2, y2 = cast(img6, True), cast(flow4, True), cast(out5, True)
    in16 = ssm_b200.flow_pack(a2.detach(), f2, td, n_timesteps=N, coord_mode=mode_name)
    frames2 = ssm_b200.fuse(a2, in16, y2, td, coord_mode=mode_name)
    # only the path through in16[:, 6:10] -> fuse: cut the warped-image channels of flow_pack out of the graph
    frames2.backward(cast(g3))
    assert torch.equal(frames, frames2), "recomputed estimated flows changed the frames"
    assert torch.equal(y.grad, y2.grad)
    assert torch.equal(a.grad, a2.grad)
    if dtype == torch.float32:
        # flow gradient of the two-step path = coefficients applied to grad in16[6:10] (+ zero through the
        # warped channels, whose upstream gradient is zero here)
        ref_gflow = []
        for n in range(N):
            tn = t[:, n]
            r16 = c_oracle.compute_inputs(img6, flow4, tn, coord_mode=mode)
            _, gx, _ = c_oracle.compute_output_image_backward(g3[:, n].contiguous(), img6, r16,
                                                              out5[:, n].contiguous(), tn, coord_mode=mode, need_img=False)
            _, gf = c_oracle.compute_inputs_backward(gx, img6, flow4, tn, coord_mode=mode, need_img=False)
            ref_gflow.append(gf)
        # both GPU paths (in-register accumulation over N, and the chain through the 16-channel tensor) against the
        # float64 sum of the oracle's per-timestep terms
        assert_sum_close(f.grad, ref_gflow, "fuse_from_flow grad flow vs C oracle")
        assert_sum_close(f2.grad, ref_gflow, "two-step grad flow vs C oracle")
    else:
        assert_close_bf16(f.grad, f2.grad.float(), "fuse_from_flow grad flow (bf16)")


def _loss_front_end_oracle(img6, flow4, out5, target, t, wts, g3, s1, s2):
    """The reference's loss front-end (losses.py:111, 152-167) and the gradients of
    L = sum_b sum_k wts[b,k] * sums[b,k] / (3HW) + sum(frames * g3), composed from the C ORACLE's warps and backward
    functions (CPU).  Returns frames [B,N,3,H,W], sums [B,2N+1] (float64), grad out5 [B,N,5,H,W] and the list of
    per-term flow4 gradients (their float64 sum is the reference)."""
    B, N = out5.shape[0], out5.shape[1]
    H, W = img6.shape[-2:]
    cnt = 3.0 * H * W
    i0, i1 = img6[:, 0:3].contiguous(), img6[:, 3:6].contiguous()
    frames = torch.empty(B, N, 3, H, W)
    sums = torch.zeros(B, 2 * N + 1, dtype=torch.float64)
    gy = torch.zeros_like(out5)
    gflow_terms = []
    for n in range(N):
        tn = t[:, n]
        y5 = out5[:, n].contiguous()
        tg = target[:, n]
        r16 = c_oracle.compute_inputs(img6, flow4, tn)
        fr = c_oracle.compute_output_image(img6, r16, y5, tn)
        frames[:, n] = fr
        sums[:, 2 * n] = (fr - tg).abs().double().flatten(1).sum(1)
        # upstream gradient of the fused frame: L1 reconstruction term + the dense term
        G = (wts[:, 2 * n].view(B, 1, 1, 1) / cnt) * torch.sign(fr - tg) + g3[:, n]
        _, gx, g5 = c_oracle.compute_output_image_backward(G.contiguous(), img6, r16, y5, tn, need_img=False)
        g16 = gx.clone()
        if s2:
            ft1 = (r16[:, 6:8] + y5[:, 1:3]).contiguous()
            ft0 = (r16[:, 8:10] + y5[:, 3:5]).contiguous()
            w0, w1 = c_oracle.warp(i0, ft0), c_oracle.warp(i1, ft1)
            sums[:, 2 * n + 1] = ((w0 - tg).abs() + (w1 - tg).abs()).double().flatten(1).sum(1)
            k = wts[:, 2 * n + 1].view(B, 1, 1, 1) / cnt
            _, d0 = c_oracle.warp_backward((k * torch.sign(w0 - tg)).contiguous(), i0, ft0, need_img=False)
            _, d1 = c_oracle.warp_backward((k * torch.sign(w1 - tg)).contiguous(), i1, ft1, need_img=False)
            g5 = g5.clone()
            g5[:, 1:3] += d1
            g5[:, 3:5] += d0
            g16[:, 6:8] += d1
            g16[:, 8:10] += d0
        gy[:, n] = g5
        _, gf = c_oracle.compute_inputs_backward(g16.contiguous(), img6, flow4, tn, need_img=False)
        gflow_terms.append(gf)
    if s1:
        f01, f10 = flow4[:, 0:2].contiguous(), flow4[:, 2:4].contiguous()
        wa, wb = c_oracle.warp(i1, f01), c_oracle.warp(i0, f10)
        sums[:, 2 * N] = ((wa - i0).abs() + (wb - i1).abs()).double().flatten(1).sum(1)
        k = wts[:, 2 * N].view(B, 1, 1, 1) / cnt
        _, da = c_oracle.warp_backward((k * torch.sign(wa - i0)).contiguous(), i1, f01, need_img=False)
        _, db = c_oracle.warp_backward((k * torch.sign(wb - i1)).contiguous(), i0, f10, need_img=False)
        gflow_terms.append(torch.cat([da, db], dim=1))
    return frames, sums, gy, gflow_terms


@pytest.mark.parametrize("s1,s2", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("B,N,H,W,kind", [(3, 1, 64, 96, "smooth"), (2, 1, 45, 77, "border"), (2, 3, 40, 72, "smooth")])
def test_fused_loss_front_end_vs_c_oracle(B, N, H, W, kind, s1, s2):
    """ssm_fuse_loss_fwd/bwd against the reference's loss front-end composed from the C oracle (losses.py:111,
    152-167): fused frames, the three L1 sums per sample, and the gradients of a weighted loss w.r.t. the stage-2
    output and the stage-1 flows."""
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=500 + H + N, kind=kind)
    target = synthetic.frames(B * N, H, W, n_frames=1, seed=77).view(B, N, 3, H, W)
    g3 = torch.randn(B, N, 3, H, W, generator=torch.Generator().manual_seed(3)) * 1e-3
    wts = torch.rand(B, 2 * N + 1, generator=torch.Generator().manual_seed(4)) + 0.5
    a, f, y = _dev(img6), _dev(flow4, True), _dev(out5, True)
    frames, sums = ssm_b200.fuse_loss(a, f, y, _dev(target), _dev(t), stage1_loss=s1, stage2_loss=s2)
    ((sums * _dev(wts)).sum() / (3 * H * W) + (frames * _dev(g3)).sum()).backward()
    r_frames, r_sums, r_gy, r_gf = _loss_front_end_oracle(img6, flow4, out5, target, t, wts, g3, s1, s2)
    assert_close_fp32(frames, r_frames, "fused-loss frames")
    # each sum has 3HW terms of O(1): the bar is relative (1e-5 of the sum's magnitude, at least 1e-5 absolute)
    rel = ((sums.detach().cpu().double() - r_sums).abs() / r_sums.abs().clamp_min(1.0)).max().item()
    assert rel <= 1e-5, "loss sums: relative error %.3e" % rel
    # sign() makes the L1 gradients discontinuous where a warped value equals its target to rounding: the oracle and the
    # kernel may pick different signs at such pixels (both are valid subgradients); they must be rare and the rest exact
    err = (y.grad.cpu() - r_gy).abs()
    k_max = (wts.max().item() / (3 * H * W))
    flips = (err > 1e-5).sum().item()
    assert flips <= max(2, err.numel() // 20000), "grad out5: %d elements beyond 1e-5 (max %.3e)" % (flips, err.max().item())
    ref64 = sum(x.double() for x in r_gf)
    errf = (f.grad.cpu().double() - ref64).abs()
    flips = (errf > 1e-5 + len(r_gf) * 2.0 ** -24 * sum(x.double().abs() for x in r_gf)).sum().item()
    assert flips <= max(2, errf.numel() // 20000), "grad flow4: %d elements beyond 1e-5 (max %.3e, loss weight %.1e)" % (
        flips, errf.max().item(), k_max)


@pytest.mark.parametrize("C", [1, 3, 5])
def test_warp_channels_and_partial_grads(C):
    B, H, W = 2, 40, 72
    gen = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, H, W, generator=gen)
    f = synthetic.flows(B, H, W, 2, flow_px=5.0, seed=9)
    g = torch.randn(B, C, H, W, generator=gen)
    want_out = c_oracle.warp(x, f)
    want_gx, want_gf = c_oracle.warp_backward(g, x, f)
    for need_x, need_f in [(True, True), (True, False), (False, True)]:
        xd, fd = _dev(x, need_x), _dev(f, need_f)
        y = ssm_b200.warp(xd, fd)
        y.backward(_dev(g))
        assert_close_fp32(y, want_out, "warp fwd")
        if need_x:
            assert_close_fp32(xd.grad, want_gx, "warp grad img")
        if need_f:
            assert_close_fp32(fd.grad, want_gf, "warp grad flow")


def test_packed_and_planar_gathers_agree_bitwise():
    """The RGBx staging copy changes how taps are fetched, not what is computed."""
    B, N, H, W = 2, 3, 40, 72
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=60, kind="border")
    a, f, y, td = _dev(img6), _dev(flow4), _dev(out5), _dev(t)
    rgbx = ssm_b200.pack_frames(a)
    assert rgbx.shape == (B, 2, H, W, 4)
    assert torch.equal(rgbx[:, 0, :, :, :3].permute(0, 3, 1, 2), a[:, 0:3])
    assert torch.equal(rgbx[:, 1, :, :, :3].permute(0, 3, 1, 2), a[:, 3:6])
    packed16 = ssm_b200.flow_pack(a, f, td, n_timesteps=N, packed=rgbx)
    for n in range(N):   # N = 1 calls take the planar path
        planar16 = ssm_b200.flow_pack(a, f, td[:, n], n_timesteps=1)
        assert torch.equal(packed16[:, n], planar16[:, 0])
        planar3 = ssm_b200.fuse(a, packed16[:, n:n + 1], y[:, n:n + 1], td[:, n])
        packed3 = ssm_b200.fuse(a, packed16[:, n:n + 1], y[:, n:n + 1], td[:, n], packed=rgbx)
        assert torch.equal(planar3, packed3)
    # bf16 storage
    ab = a.bfloat16()
    rb = ssm_b200.pack_frames(ab)
    assert torch.equal(rb[:, 1, :, :, :3].permute(0, 3, 1, 2), ab[:, 3:6])
    p16 = ssm_b200.flow_pack(ab, f.bfloat16(), td, n_timesteps=N, packed=rb)
    q16 = ssm_b200.flow_pack(ab, f.bfloat16(), td[:, 1], n_timesteps=1)
    assert torch.equal(p16[:, 1], q16[:, 0])


def test_strided_views_are_accepted():
    """compute_output_image receives channel-sliced views (flow_interpolation.py:402-403)."""
    B, H, W = 2, 32, 64
    img6, flow4, out5, t = _inputs(B, 1, H, W, seed=77)
    big = torch.randn(B, 9, H, W)
    big[:, 2:8] = img6
    view = _dev(big)[:, 2:8]                       # non-contiguous batch stride
    in16 = ssm_b200.flow_pack(view, _dev(flow4), _dev(t), n_timesteps=1)
    ref16 = c_oracle.compute_inputs(img6, flow4, t[:, 0])
    assert_close_fp32(in16[:, 0], ref16, "flow_pack on a channel-sliced view")
    y = ssm_b200.warp(view[:, 0:3], _dev(flow4)[:, 2:4])
    assert_close_fp32(y, c_oracle.warp(img6[:, 0:3].contiguous(), flow4[:, 2:4].contiguous()), "warp on views")


# ---------------------------------------------------------------------------------------------
def test_against_reference_ops_on_cuda():
    """The reference's own torch ops on the CUDA device.  cuDNN off = ATen's CUDA kernels, matched
    by coord mode "cuda" at 1e-5 (fwd and grads); the CPU-vs-CUDA and cuDNN deltas of the REFERENCE
    are printed for context (they exceed 1e-5 on their own, SURVEY.md findings 3/3b)."""
    B, H, W = 2, 352, 352
    img6, flow4, out5, t = _inputs(B, 1, H, W, seed=300, flow_px=10.0)
    t4 = t.view(B, 1, 1, 1)
    g16 = torch.randn(B, 16, H, W, generator=torch.Generator().manual_seed(1))
    g3 = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(2))

    def run_ref(cudnn):
        prev = torch.backends.cudnn.enabled
        torch.backends.cudnn.enabled = cudnn
        try:
            a, b = _dev(img6, True), _dev(flow4, True)
            in16 = torch_oracle.compute_inputs(a, b, _dev(t4))
            in16.backward(_dev(g16))
            a2, xin, yo = _dev(img6, True), in16.detach().clone().requires_grad_(True), _dev(out5[:, 0], True)
            fr = torch_oracle.compute_output_image(a2, xin, yo, _dev(t4))
            fr.backward(_dev(g3))
            return dict(in16=in16.detach(), gflow=b.grad, gimg=a.grad, frame=fr.detach(), gin=xin.grad, gout=yo.grad,
                        gimg2=a2.grad)
        finally:
            torch.backends.cudnn.enabled = prev

    def run_new(mode):
        a, b = _dev(img6, True), _dev(flow4, True)
        in16 = ssm_b200.flow_pack(a, b, _dev(t), n_timesteps=1, coord_mode=mode)
        in16.backward(_dev(g16).unsqueeze(1))
        a2, xin, yo = _dev(img6, True), in16.detach().clone().requires_grad_(True), _dev(out5, True)
        fr = ssm_b200.fuse(a2, xin, yo, _dev(t), coord_mode=mode)
        fr.backward(_dev(g3).unsqueeze(1))
        return dict(in16=in16.detach()[:, 0], gflow=b.grad, gimg=a.grad, frame=fr.detach()[:, 0], gin=xin.grad[:, 0],
                    gout=yo.grad[:, 0], gimg2=a2.grad)

    aten = run_ref(cudnn=False)
    cudnn = run_ref(cudnn=True)
    new_cuda = run_new("cuda")
    new_cpu = run_new("cpu")
    report = {k: (max_err(new_cuda[k], aten[k]), max_err(new_cpu[k], aten[k]), max_err(cudnn[k], aten[k])) for k in aten}
    print("\nmax abs diff vs reference-on-CUDA (ATen): key: new[mode=cuda]  new[mode=cpu]  reference-cuDNN")
    for k, v in report.items():
        print("  %-6s %.2e  %.2e  %.2e" % ((k,) + v))
    for k in aten:
        assert_close_fp32(new_cuda[k], aten[k], "mode=cuda vs reference ATen-CUDA: " + k)


def test_speed_against_reference_ops_on_cuda():
    """Same box, same device: the reference's per-timestep torch-op path (what the reference runs on a
    GPU: ~110 ATen launches per timestep, cuDNN grid sampler) vs the three launches of the batched path,
    at BASELINE configs[0] size (1 pair 736x1280, 7 timesteps) and for 4 pairs of 1088x1920.  Context
    only (SURVEY.md section 8(d) last row); asserts just that the new path is not slower."""
    import time
    print()
    for B, H, W in ((1, 736, 1280), (4, 1088, 1920)):
        N = 7
        img6 = synthetic.frames(B, H, W, seed=42, device=DEV)
        flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=DEV)
        out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=DEV)
        t = synthetic.timesteps(B, N, device=DEV)

        def ref():
            outs = []
            for n in range(N):
                tn = t[:, n].view(B, 1, 1, 1)
                in16 = torch_oracle.compute_inputs(img6, flow4, tn)
                outs.append(torch_oracle.compute_output_image(img6, in16, out5[:, n], tn))
            return outs

        def new():
            rgbx = ssm_b200.pack_frames(img6)
            in16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=rgbx)
            return in16, ssm_b200.fuse_from_flow(img6, flow4, out5, t, packed=rgbx)

        times = {}
        with torch.no_grad():
            for name, fn in (("reference torch ops (cuDNN sampler)", ref), ("ssm_b200", new)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
                times[name] = (time.perf_counter() - t0) / 10
        r, n_ = times["reference torch ops (cuDNN sampler)"], times["ssm_b200"]
        print("  %d pair(s) %dx%d x %d timesteps: reference ops on CUDA %.2f ms (%.0f frames/s), ssm_b200 %.2f ms "
              "(%.0f frames/s): %.1fx" % (B, H, W, N, 1e3 * r, B * N / r, 1e3 * n_, B * N / n_, r / n_))
        assert n_ < r


# ---------------------------------------------------------------------------------------------
def test_image_gradient_is_deterministic():
    """Bit-identical run to run (the reference's atomicAdd scatter is not)."""
    B, N, H, W = 2, 7, 96, 128
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=500, kind="noise", smooth=False)
    flow4 = flow4 * 0.0 + flow4.mean(dim=(2, 3), keepdim=True)   # constant flow: many sources per destination
    g3 = torch.randn(B, N, 3, H, W, generator=torch.Generator().manual_seed(3))
    outs = []
    for _ in range(3):
        a, b = _dev(img6, True), _dev(flow4, True)
        in16 = ssm_b200.flow_pack(a, b, _dev(t), n_timesteps=N)
        fr = ssm_b200.fuse(a, in16, _dev(out5), _dev(t))
        fr.backward(_dev(g3))
        outs.append((a.grad.clone(), b.grad.clone()))
    for gi, gf in outs[1:]:
        assert torch.equal(gi, outs[0][0]) and torch.equal(gf, outs[0][1])


@pytest.mark.parametrize("mode_name,mode", MODES)
def test_bf16_storage(mode_name, mode):
    """bf16 storage, fp32 math: against the fp32 oracle on the same bf16-rounded inputs."""
    B, N, H, W = 2, 2, 64, 96
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=700, flow_px=3.0)
    img6, flow4, out5 = (x.bfloat16() for x in (img6, flow4, out5))
    g3 = (torch.randn(B, N, 3, H, W, generator=torch.Generator().manual_seed(4)) * 0.25).bfloat16()
    a, b = _dev(img6), _dev(flow4, True)
    in16 = ssm_b200.flow_pack(a, b, _dev(t), n_timesteps=N, coord_mode=mode_name)
    assert in16.dtype == torch.bfloat16
    xin, yo = in16.detach().clone().requires_grad_(True), _dev(out5, True)
    fr = ssm_b200.fuse(a, xin, yo, _dev(t), coord_mode=mode_name)
    fr.backward(_dev(g3))
    for n in range(N):
        tn = t[:, n]
        r16 = c_oracle.compute_inputs(img6.float(), flow4.float(), tn, coord_mode=mode)
        assert_close_bf16(in16[:, n], r16, "bf16 flow_pack fwd")
        x16 = in16[:, n].detach().float().cpu()
        r3 = c_oracle.compute_output_image(img6.float(), x16, out5[:, n].float().contiguous(), tn, coord_mode=mode)
        assert_close_bf16(fr[:, n], r3, "bf16 fuse fwd")
        _, gx, gy = c_oracle.compute_output_image_backward(g3[:, n].float().contiguous(), img6.float(), x16,
                                                           out5[:, n].float().contiguous(), tn, coord_mode=mode,
                                                           need_img=False)
        assert_close_bf16(yo.grad[:, n], gy, "bf16 fuse grad out5")
        assert_close_bf16(xin.grad[:, n], gx, "bf16 fuse grad in16")


# ---------------------------------------------------------------------------------------------
def test_full_size_properties():
    """BASELINE.json configs[1]: 16 pairs 1088x1920, 7 timesteps, on one B200.  Size-independent
    properties on the whole batch + the C oracle on two (pair, timestep) slices."""
    B, N, H, W = 16, 7, 1088, 1920
    img6 = synthetic.frames(B, H, W, seed=42, device=DEV)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=DEV)
    t = synthetic.timesteps(B, N, device=DEV)
    in16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N)
    frames = ssm_b200.fuse(img6, in16, out5, t)
    assert torch.isfinite(frames).all()
    # layout identities (flow_interpolation.py:364-367), exact
    for n in (0, N - 1):
        assert torch.equal(in16[:, n, 0:3], img6[:, 3:6]) and torch.equal(in16[:, n, 13:16], img6[:, 0:3])
    # warped channels == stand-alone warp with the packed flows
    w1 = ssm_b200.warp(img6[:, 3:6], in16[:, 3, 6:8])
    assert max_err(w1, in16[:, 3, 3:6]) <= 1e-6
    # idempotence / determinism of the forward
    assert torch.equal(ssm_b200.fuse(img6, in16, out5, t), frames)
    # time-reversal symmetry: swap frames, flows, residuals, negate the logit, t -> 1 - t
    b = 5
    img_sw = torch.cat([img6[b:b + 1, 3:6], img6[b:b + 1, 0:3]], 1)
    in_sw = in16[b:b + 1].clone()
    in_sw[:, :, 6:8], in_sw[:, :, 8:10] = in16[b:b + 1, :, 8:10], in16[b:b + 1, :, 6:8]
    out_sw = torch.cat([-out5[b:b + 1, :, 0:1], out5[b:b + 1, :, 3:5], out5[b:b + 1, :, 1:3]], 2)
    fr_sw = ssm_b200.fuse(img_sw, in_sw, out_sw, 1.0 - t[b:b + 1])
    assert max_err(fr_sw, frames[b:b + 1]) <= 5e-6
    # two slices against the C oracle
    for (b, n) in ((0, 0), (15, 6)):
        i6, f4 = img6[b:b + 1].cpu(), flow4[b:b + 1].cpu()
        tn = t[b:b + 1, n].cpu()
        r16 = c_oracle.compute_inputs(i6, f4, tn)
        assert_close_fp32(in16[b:b + 1, n], r16, "full-size flow_pack (%d,%d)" % (b, n))
        r3 = c_oracle.compute_output_image(i6, r16, out5[b:b + 1, n].cpu().contiguous(), tn)
        assert_close_fp32(frames[b:b + 1, n], r3, "full-size fuse (%d,%d)" % (b, n))


def test_c5_4k_31_timesteps_properties():
    """BASELINE.json configs[4]: one 2176x3840 pair, 31 intermediate times (k/32): the (pair, timestep)
    shards of 8 ranks (4/4/4/4/4/4/4/3) reproduce the single-launch result bit for bit, the recomputed-flow
    kernel equals the two-step path, and one timestep is checked against the C oracle."""
    from ssm_b200 import sharding
    B, N, H, W = 1, 31, 2176, 3840
    img6 = synthetic.frames(B, H, W, seed=42, device=DEV)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=43, device=DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=DEV)
    t = synthetic.timesteps(B, N, device=DEV)
    assert abs(t[0, 0].item() - 1 / 32) < 1e-7 and abs(t[0, 30].item() - 31 / 32) < 1e-7
    in16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N)
    frames = ssm_b200.fuse_from_flow(img6, flow4, out5, t)
    assert torch.isfinite(frames).all()
    assert torch.equal(ssm_b200.fuse(img6, in16, out5, t), frames)
    counts = []
    for rank in range(8):
        (p, t0, t1), = sharding.shard_work(1, N, rank, 8)
        counts.append(t1 - t0)
        part = ssm_b200.fuse_from_flow(img6, flow4, out5[:, t0:t1].contiguous(), t[:, t0:t1].contiguous())
        assert torch.equal(part, frames[:, t0:t1]), "rank %d shard differs" % rank
    assert counts == [4, 4, 4, 4, 4, 4, 4, 3]
    n = 17
    i6, f4, tn = img6.cpu(), flow4.cpu(), t[:, n].cpu()
    r16 = c_oracle.compute_inputs(i6, f4, tn)
    assert_close_fp32(in16[:, n], r16, "4K flow_pack n=17")
    assert_close_fp32(frames[:, n], c_oracle.compute_output_image(i6, r16, out5[:, n].cpu().contiguous(), tn), "4K fuse n=17")


def test_c3_training_backward_is_linear_in_the_upstream_gradient():
    """BASELINE.json configs[2] shape (64 crops of 352x352, per-sample random t): the backward of the
    path is a linear map of the upstream gradient -- bwd(2 g1 - 3 g2) = 2 bwd(g1) - 3 bwd(g2) -- which
    checks every gradient kernel at full batch size without an oracle run; one sample is also checked
    against the C oracle."""
    B, N, H, W = 64, 1, 352, 352
    img6 = synthetic.frames(B, H, W, seed=42, device=DEV)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=8.0, seed=43, device=DEV).requires_grad_(True)
    out5 = synthetic.unet_out5(B, N, H, W, seed=44, device=DEV).requires_grad_(True)
    t = synthetic.random_timesteps(B, 1, seed=45).to(DEV)
    in16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=1)
    frames = ssm_b200.fuse_from_flow(img6, flow4, out5, t)
    gen = torch.Generator(device=DEV).manual_seed(1)
    g1, g2 = torch.randn(frames.shape, device=DEV, generator=gen), torch.randn(frames.shape, device=DEV, generator=gen)
    h1, h2 = torch.randn(in16.shape, device=DEV, generator=gen), torch.randn(in16.shape, device=DEV, generator=gen)

    def bwd(g, h):
        return torch.autograd.grad([frames, in16], [flow4, out5], [g, h], retain_graph=True)

    a, b, c = bwd(g1, h1), bwd(g2, h2), bwd(2 * g1 - 3 * g2, 2 * h1 - 3 * h2)
    for x, y, z, what in ((a[0], b[0], c[0], "flow"), (a[1], b[1], c[1], "out5")):
        want = 2 * x - 3 * y
        assert max_err(z, want) <= 1e-5 * max(1.0, want.abs().max().item()), "backward not linear in grad (%s)" % what
    s = 37
    i6, f4, y5, ts = img6[s:s + 1].cpu(), flow4[s:s + 1].detach().cpu(), out5[s:s + 1, 0].detach().cpu(), t[s:s + 1, 0].cpu()
    r16 = c_oracle.compute_inputs(i6, f4, ts)
    _, gx, gy = c_oracle.compute_output_image_backward(g1[s:s + 1, 0].cpu(), i6, r16, y5, ts, need_img=False)
    # autograd adds the flow gradients of the two consumers (frames through in16[6:10], and in16 itself) in fp32:
    # the oracle supplies the two terms separately
    _, gf_a = c_oracle.compute_inputs_backward(gx, i6, f4, ts, need_img=False)
    _, gf_b = c_oracle.compute_inputs_backward(h1[s:s + 1, 0].cpu().contiguous(), i6, f4, ts, need_img=False)
    assert_close_fp32(a[1][s:s + 1, 0], gy, "C3 grad out5, sample 37")
    assert_sum_close(a[0][s:s + 1], [gf_a, gf_b], "C3 grad flow, sample 37")


def test_streams_threads_and_argument_errors():
    """Boundary behaviour: calls enqueue on the caller's current stream and
    never synchronise; concurrent Python threads (nn.DataParallel style) are safe; bad arguments raise."""
    import threading
    B, N, H, W = 2, 3, 64, 96
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=700)
    a, f, y, td = _dev(img6), _dev(flow4), _dev(out5), _dev(t)
    want = ssm_b200.fuse_from_flow(a, f, y, td)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        got = ssm_b200.fuse_from_flow(a, f, y, td)
    side.synchronize()
    assert torch.equal(got, want)
    results = [None] * 4

    def worker(i):
        with torch.cuda.stream(torch.cuda.Stream()):
            r = ssm_b200.fuse(a, ssm_b200.flow_pack(a, f, td, n_timesteps=N), y, td)
            torch.cuda.current_stream().synchronize()
            results[i] = r
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert all(torch.equal(r, want) for r in results)
    with pytest.raises(RuntimeError):
        ssm_b200.flow_pack(a, f[:, :3], td, n_timesteps=N)                  # wrong channel count
    with pytest.raises(RuntimeError):
        ssm_b200.fuse_from_flow(a, f, y[:, :, :4], td)
    with pytest.raises(RuntimeError):
        ssm_b200.flow_pack(a, f.to(torch.bfloat16), td, n_timesteps=N)      # mixed storage types
    with pytest.raises(RuntimeError):
        ssm_b200.flow_pack(a, f, td[:, :2], n_timesteps=N)                  # t count
    with pytest.raises(AssertionError):
        ssm_b200.SynthesisMixin().compute_inputs(a, f, torch.full((B, 1, 1, 1), 1.0))   # validators.py:9-11
    # non-finite flows are sampled outside the image (zeros), finite everywhere else
    f_bad = f.clone()
    f_bad[0, :, 5, 7] = float("inf")
    out = ssm_b200.flow_pack(a, f_bad, td, n_timesteps=N)
    assert torch.isfinite(out[:, :, 3:6]).all() and torch.isfinite(out[:, :, 10:13]).all()


def test_caller_owned_output_buffers():
    """out= (caller-owned result buffers, inference): same bits as the allocating call, results land in
    the given storage, and mismatched buffers / gradient-requiring inputs are rejected."""
    B, N, H, W = 2, 3, 64, 96
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=1200)
    a, f, y, td = _dev(img6), _dev(flow4), _dev(out5), _dev(t)
    rgbx = torch.empty((B, 2, H, W, 4), device=DEV)
    b16 = torch.full((B, N, 16, H, W), float("nan"), device=DEV)
    b3 = torch.full((B, N, 3, H, W), float("nan"), device=DEV)
    b3b = torch.full((B, N, 3, H, W), float("nan"), device=DEV)
    r = ssm_b200.pack_frames(a, out=rgbx)
    assert r.data_ptr() == rgbx.data_ptr() and torch.equal(rgbx, ssm_b200.pack_frames(a))
    w16 = ssm_b200.flow_pack(a, f, td, n_timesteps=N, packed=rgbx, out=b16)
    assert w16.data_ptr() == b16.data_ptr() and torch.equal(b16, ssm_b200.flow_pack(a, f, td, n_timesteps=N))
    w3 = ssm_b200.fuse_from_flow(a, f, y, td, packed=rgbx, out=b3)
    assert w3.data_ptr() == b3.data_ptr() and torch.equal(b3, ssm_b200.fuse_from_flow(a, f, y, td))
    ssm_b200.fuse(a, b16, y, td, out=b3b)
    assert torch.equal(b3b, b3)
    with pytest.raises(RuntimeError):
        ssm_b200.flow_pack(a, f, td, n_timesteps=N, out=b16[:, :2])                       # wrong shape
    with pytest.raises(RuntimeError):
        ssm_b200.fuse_from_flow(a, f, y, td, out=b3.bfloat16())                           # wrong dtype
    with pytest.raises(RuntimeError):
        ssm_b200.fuse_from_flow(a, _dev(flow4, True), y, td, out=b3)                      # needs autograd
    with torch.no_grad():                                                                 # fine without it
        ssm_b200.fuse_from_flow(a, _dev(flow4, True), y, td, out=b3)


def test_host_entry_point_matches_device_path():
    """ssm_synthesize_host (host buffers, copies inside) == device path."""
    B, N, H, W = 4, 3, 64, 96
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=900)
    out3, in16 = ssm_b200.synthesize_host(img6.pin_memory(), flow4.pin_memory(), out5.pin_memory(), t,
                                          return_inputs=True)
    d16 = ssm_b200.flow_pack(_dev(img6), _dev(flow4), _dev(t), n_timesteps=N)
    d3 = ssm_b200.fuse(_dev(img6), d16, _dev(out5), _dev(t))
    assert torch.equal(in16, d16.cpu()) and torch.equal(out3, d3.cpu())
    for n in range(N):
        r16 = c_oracle.compute_inputs(img6, flow4, t[:, n])
        assert_close_fp32(out3[:, n], c_oracle.compute_output_image(img6, r16, out5[:, n].contiguous(), t[:, n]),
                          "host entry point vs oracle")


@pytest.mark.parametrize("size", [1920, 1088, 352, 1280, 736, 3840, 2176, 2, 1, 37, 4097])
def test_constant_divisor_division_is_ieee_exact(size):
    """Known-answer self-test: the kernels' 5-instruction division by max(size-1,1) gives the same
    normalised coordinate rn(s/d - 1) as the IEEE division for every finite fp32 dividend (2^32
    cases per divisor).  Raw quotients may differ only in the sign of zero / denormal range."""
    import ctypes
    from ssm_b200 import _abi
    counter = torch.zeros(3, dtype=torch.int64, device=DEV)
    rc = _abi.lib().ssm_selftest_division(size, ctypes.c_void_p(counter.data_ptr()), _abi.stream_ptr(torch.device(DEV)))
    _abi.check(rc, "ssm_selftest_division")
    torch.cuda.synchronize()
    bad_n, bad_q, example = counter.tolist()
    print("size %d: coordinate mismatches %d, raw-quotient mismatches %d (example dividend bits 0x%08x)"
          % (size, bad_n, bad_q, example))
    assert bad_n == 0, "%d coordinate mismatches for size %d" % (bad_n, size)
    assert bad_q <= 2 ** 25, "raw quotient mismatches beyond the denormal range for size %d" % size


# ---------------------------------------------------------------------------------------------
def test_large_flow_reference_fixture():
    """352 x 352 with flows of up to 82 px: the CUDA path against outputs and autograd gradients of the reference itself
    (tests/golden/make_golden_large.py), 1e-5, on the stored band of rows -- fp32 frames (RGBx gathers) and the 8-bit-frame
    kernels (the fixture's frames are uint8 images normalised with the reference's expression)."""
    from ssm_b200 import q8
    from util import load_large_golden
    d = load_large_golden()
    rows = d["rows"]
    B = d["img6"].shape[0]
    t = d["t"].view(B, 1)
    a, f = _dev(d["img6"]), _dev(d["flow4"], True)
    in16 = ssm_b200.flow_pack(a, f, _dev(t), n_timesteps=1)
    in16.backward(_dev(d["g16"]).unsqueeze(1))
    warped = torch.cat([in16[:, 0, 3:6], in16[:, 0, 10:13]], 1)[:, :, rows]
    assert_close_fp32(warped, d["in16_warped"], "compute_inputs warped images, large flows")
    assert_close_fp32(f.grad[:, :, rows], d["pack_gflow"], "compute_inputs grad flow, large flows")
    xin, yo = _dev(in16.detach()[:, 0], True), _dev(d["out5"], True)
    frame = ssm_b200.SynthesisMixin().compute_output_image(a, xin, yo, _dev(t.view(B, 1, 1, 1)))
    frame.backward(_dev(d["g3"]))
    assert_close_fp32(frame[:, :, rows], d["frame"], "compute_output_image, large flows")
    assert_close_fp32(xin.grad[:, 6:10, rows], d["fuse_gflows"], "compute_output_image grad in16[6:10], large flows")
    assert_close_fp32(yo.grad[:, :, rows], d["fuse_gout5"], "compute_output_image grad out5, large flows")
    # the 8-bit-frame kernels on the same uint8 images
    planar, quads, norm, _ = q8.prepare(d["u8"].to(DEV), order="rgb", lut=ssm_b200.normalisation_lut(device="cpu"))
    H, W = planar.shape[-2:]
    img6_q = planar.view(B, 6, H, W)
    assert torch.equal(img6_q.cpu(), d["img6"]), "normalised frames are not bit-identical to the reference's"
    in16_q = q8.flow_pack(img6_q, quads, _dev(d["flow4"]), _dev(t), norm, n_timesteps=1)
    warped_q = torch.cat([in16_q[:, 0, 3:6], in16_q[:, 0, 10:13]], 1)[:, :, rows]
    assert_close_fp32(warped_q, d["in16_warped"], "compute_inputs from uint8 frames, large flows")
    frame_q = q8.fuse_from_flow(quads, _dev(d["flow4"]), _dev(d["out5"]).unsqueeze(1), _dev(t), norm)
    assert_close_fp32(frame_q[:, 0, :, rows], d["frame"], "compute_output_image from uint8 frames, large flows")


def test_patch_reference_on_a_reference_style_module():
    """patch_reference as INTEGRATION.md section 1 uses it, on a stand-in for the reference's `models` modules (the
    reference tree does not travel to the GPU box): a FlowInterpolationModel class whose three methods and whose
    module-level warp are the reference's op sequence (oracle/torch_oracle.py).  After patching, CUDA tensors run the
    B200 path and reproduce the reference's golden outputs and gradients; CPU tensors still reach the stand-in's own code."""
    import types

    class FlowInterpolationModel:
        verbose = False

        def compute_inputs(self, img_tensor, flow_pred_tensor, t):
            return torch_oracle.compute_inputs(img_tensor, flow_pred_tensor, t)

        def extract_outputs(self, output_tensor):
            return torch_oracle.extract_outputs(output_tensor)

        def compute_output_image(self, img_tensor, input_tensor, output_tensor, t):
            return torch_oracle.compute_output_image(img_tensor, input_tensor, output_tensor, t)

    fi = types.SimpleNamespace(FlowInterpolationModel=FlowInterpolationModel, warp=torch_oracle.warp)
    layers = types.SimpleNamespace(warp=torch_oracle.warp)
    losses = types.SimpleNamespace(warp=torch_oracle.warp)
    ssm_b200.patch_reference(fi, layers, losses)
    model = fi.FlowInterpolationModel()
    for name in ("small_smooth", "border"):
        d = load_golden(name)
        B = d["img6"].shape[0]
        t = d["t"].view(B, 1, 1, 1)
        a, b = _dev(d["img6"], True), _dev(d["flow4"], True)
        in16 = model.compute_inputs(a, b, _dev(t))
        in16.backward(_dev(d["pack_g16"]))
        assert_close_fp32(in16, d["in16"], "patched compute_inputs")
        assert_close_fp32(b.grad, d["pack_gflow"], "patched compute_inputs grad flow")
        xin, yo = _dev(d["in16"], True), _dev(d["out5"], True)
        frame = model.compute_output_image(_dev(d["img6"]), xin, yo, _dev(t))
        frame.backward(_dev(d["fuse_g3"]))
        assert_close_fp32(frame, d["frame"], "patched compute_output_image")
        assert_close_fp32(yo.grad, d["fuse_gout5"], "patched compute_output_image grad out5")
        v1, df1, df0, v0 = model.extract_outputs(_dev(d["out5"]))
        assert_close_fp32(v1 + v0, torch.ones_like(v1), "patched extract_outputs")
        x, f = _dev(d["img6"][:, 0:3], True), _dev(d["flow4"][:, 0:2], True)
        y = losses.warp(x, f)                                  # losses.py:152-161 calls the module-level name
        y.backward(_dev(d["warp_gout"]))
        assert_close_fp32(y, d["warp_out"], "patched losses.warp")
        assert_close_fp32(x.grad, d["warp_gimg"], "patched losses.warp grad img")
        assert_close_fp32(f.grad, d["warp_gflow"], "patched losses.warp grad flow")
        assert torch.equal(layers.warp(d["img6"][:, 0:3], d["flow4"][:, 0:2]), d["warp_out"]), "CPU tensors: the module's own warp"
        assert torch.equal(model.compute_inputs(d["img6"], d["flow4"], t), d["in16"])
    # autocast hands over a bf16 U-Net output next to fp32 frames: promoted, as the reference's torch ops do
    d = load_golden("small_smooth")
    t = _dev(d["t"].view(-1, 1, 1, 1))
    mixed = model.compute_output_image(_dev(d["img6"]), _dev(d["in16"]), _dev(d["out5"]).bfloat16(), t)
    want = model.compute_output_image(_dev(d["img6"]), _dev(d["in16"]), _dev(d["out5"]).bfloat16().float(), t)
    assert mixed.dtype == torch.float32 and torch.equal(mixed, want)


def test_device_side_t_is_range_checked_on_request():
    """validators.py:9-11 asserts 0 < t < 1.  A device-side t is checked without a synchronisation in "flag" mode."""
    B, N, H, W = 2, 3, 16, 32
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=77)
    previous = ssm_b200.set_device_t_check("flag")
    try:
        ssm_b200.t_violations()
        ssm_b200.flow_pack(_dev(img6), _dev(flow4), _dev(t), n_timesteps=N)
        assert ssm_b200.t_violations() == 0
        bad = t.clone()
        bad[0, 1], bad[1, 2] = 1.0, -0.25
        ssm_b200.flow_pack(_dev(img6), _dev(flow4), _dev(bad), n_timesteps=N)
        assert ssm_b200.t_violations() == 2
        assert ssm_b200.t_violations() == 0            # reading resets the counter
    finally:
        ssm_b200.set_device_t_check(previous)
    with pytest.raises(AssertionError):                # a host-side t is always checked
        ssm_b200.flow_pack(_dev(img6), _dev(flow4), bad, n_timesteps=N)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_warp_from_packed_image_is_bit_identical(dtype):
    """warp(x, flo, packed=pack_image(x)): the RGBx gathers read the same values as the planar ones, so output, flow
    gradient and (deterministic) image gradient are the same bits; one packed copy serves several warps of an image."""
    B, H, W = 2, 45, 77
    img6, flow4, _, _ = _inputs(B, 1, H, W, seed=4100, kind="border")
    g = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(9)).to(DEV).to(dtype)
    x1, x2 = _dev(img6[:, 0:3].to(dtype), True), _dev(img6[:, 0:3].to(dtype), True)
    packed = ssm_b200.pack_image(x2)
    assert packed.shape == (B, H, W, 4) and torch.equal(packed[..., :3].permute(0, 3, 1, 2), x2.detach())
    for k in (0, 2):                                  # two different flows, one packed copy
        f1, f2 = _dev(flow4[:, k:k + 2].to(dtype), True), _dev(flow4[:, k:k + 2].to(dtype), True)
        y1 = ssm_b200.warp(x1, f1)
        y2 = ssm_b200.warp(x2, f2, packed=packed)
        assert torch.equal(y1, y2)
        x1.grad = x2.grad = None
        y1.backward(g)
        y2.backward(g)
        assert torch.equal(f1.grad, f2.grad) and torch.equal(x1.grad, x2.grad)
    with pytest.raises(RuntimeError):
        ssm_b200.warp(x1, _dev(flow4[:, 0:2].to(dtype)), packed=packed[:, :, :, :3].contiguous())


@pytest.mark.parametrize("H,W", [(32, 64), (45, 150)])
def test_image_gradient_when_every_pixel_lands_on_one_cell(H, W):
    """A flow that collapses the whole image onto one pixel: that destination receives H*W maximal contributions, so
    the int32 cells of the shared-memory windows wrap many times -- the overflow repair of the segmented scatter
    (csrc/ssm_scatter.cuh) must keep the sum exact.  Against the analytic value and the C oracle."""
    cy, cx = H // 2 + 1, W // 3
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    flo = torch.stack([cx - xs, cy - ys]).unsqueeze(0).contiguous()             # integer displacements: weight 1 on one tap
    img = synthetic.frames(1, H, W, seed=5)[:, 0:3].contiguous()
    gout = torch.full((1, 3, H, W), 1.5)
    gout[0, 1] = -0.75
    gout[0, 2, ::2] = 2.0                                                       # channel 2: rows alternate 2.0 / 1.5
    x, f = _dev(img, True), _dev(flo)
    y = ssm_b200.warp(x, f)
    y.backward(_dev(gout))
    gi_ref, _ = c_oracle.warp_backward(gout, img, flo)
    # the reference's coordinate round trip is not the identity: for some sizes a little weight leaks to the
    # neighbouring cells, so the oracle -- not the analytic single-cell value -- is the yardstick.  The oracle sums H*W
    # fp32 terms one after the other; the kernel's integer sum is exact, so allow the oracle its summation error.
    total = gout.sum(dim=(2, 3))[0]
    assert (gi_ref[0, :, cy - 1:cy + 2, cx - 1:cx + 2].sum(dim=(1, 2)) - total).abs().max() <= 1e-3 * total.abs().max()
    # (H*W roundings of a running sum of magnitude |total|: measured 3.7e-6 relative at 45 x 150)
    err = (x.grad.cpu() - gi_ref).abs().max().item()
    assert err <= 1e-5 + 2e-5 * total.abs().max().item(), "collapsed image gradient: max abs err %.3e" % err
    if (H, W) == (32, 64):      # here the round trip is exact: everything lands on the one cell, and the sum is exact
        want = torch.zeros(1, 3, H, W)
        want[0, :, cy, cx] = total
        assert torch.equal(x.grad.cpu(), want)
    again = _dev(img, True)
    ssm_b200.warp(again, f).backward(_dev(gout))
    assert torch.equal(again.grad, x.grad)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flow_px,uniform,N", [(6.0, False, 3), (90.0, False, 3), (6.0, True, 7)])
def test_image_gradient_windows_across_timesteps(flow_px, uniform, N):
    """The image-gradient windows of csrc/ssm_scatter.cuh are kept across the timesteps of a frame and re-anchored only
    when the tile's centre displacement has drifted more than 12 px.  6 px flows: one window per frame serves all N
    timesteps; 90 px flows: the displacement moves ~22 px per timestep at N = 3, so the window is flushed and re-anchored
    between timesteps; a UNIFORM 6 px flow with a constant upstream gradient at N = 7 piles seven timesteps of same-sign,
    near-maximal contributions on the cells of a kept window, which wraps its int32 cells (7 x 1.5 x 2^28 > 2^31): the
    overflow repair must keep the sums exact.  Image gradients of compute_inputs and compute_output_image against the
    float64 sum of the C oracle's per-timestep terms; run twice for bit identity."""
    B, H, W = 1, 96, 256
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=4242, kind="smooth", flow_px=flow_px)
    gen = torch.Generator().manual_seed(11)
    g16 = torch.randn(B, N, 16, H, W, generator=gen)
    g3 = torch.randn(B, N, 3, H, W, generator=gen)
    if uniform:
        flow4 = torch.zeros_like(flow4)
        flow4[:, 0], flow4[:, 2] = flow_px, -flow_px      # F_0->1 = +6 px, F_1->0 = -6 px horizontally
        out5 = out5 * 0.0
        g16, g3 = torch.full_like(g16, 3.0), torch.full_like(g3, 3.0)
    grads = []
    for _ in range(2):
        a, b = _dev(img6, True), _dev(flow4, True)
        ssm_b200.flow_pack(a, b, _dev(t), n_timesteps=N).backward(_dev(g16))
        a2, b2, yo = _dev(img6, True), _dev(flow4, True), _dev(out5, True)
        ssm_b200.fuse_from_flow(a2, b2, yo, _dev(t)).backward(_dev(g3))
        grads.append((a.grad.clone(), a2.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    ref1, ref2 = [], []
    for n in range(N):
        tn = t[:, n]
        r16 = c_oracle.compute_inputs(img6, flow4, tn)
        ref1.append(c_oracle.compute_inputs_backward(g16[:, n].contiguous(), img6, flow4, tn)[0])
        ref2.append(c_oracle.compute_output_image_backward(g3[:, n].contiguous(), img6, r16, out5[:, n].contiguous(), tn)[0])
    assert_sum_close(grads[0][0], ref1, "compute_inputs grad img (flow %g px)" % flow_px)
    assert_sum_close(grads[0][1], ref2, "compute_output_image grad img (flow %g px)" % flow_px)


# ---------------------------------------------------------------------------------------------
def test_empty_batch_matches_the_reference_ops():
    """A 0 x C x H x W batch: the reference's torch ops return empty tensors with empty gradients (checked against the
    reference itself on CPU in the build container and, here, against its restated ops on the device); nothing is
    launched, shapes and gradient shapes are the reference's."""
    H, W = 8, 12
    x = torch.zeros((0, 3, H, W), device=DEV, requires_grad=True)
    f = torch.zeros((0, 2, H, W), device=DEV, requires_grad=True)
    y = ssm_b200.warp(x, f)
    ref = torch_oracle.warp(x.detach(), f.detach())
    assert y.shape == ref.shape == (0, 3, H, W)
    y.sum().backward()
    assert x.grad.shape == x.shape and f.grad.shape == f.shape
    ops = ssm_b200.SynthesisMixin()
    img = torch.zeros((0, 6, H, W), device=DEV, requires_grad=True)
    flow = torch.zeros((0, 4, H, W), device=DEV, requires_grad=True)
    t = torch.zeros((0, 1, 1, 1), device=DEV)
    in16 = ops.compute_inputs(img, flow, t)
    assert in16.shape == torch_oracle.compute_inputs(img.detach(), flow.detach(), t).shape == (0, 16, H, W)
    out5 = torch.zeros((0, 5, H, W), device=DEV, requires_grad=True)
    frame = ops.compute_output_image(img, in16, out5, t)
    assert frame.shape == (0, 3, H, W)
    frame.sum().backward()
    assert flow.grad.shape == flow.shape and out5.grad.shape == out5.shape and img.grad.shape == img.shape
    # timestep-batched forms: no pairs, or no timesteps
    assert ops.compute_inputs_batched(img, flow, torch.zeros((0, 7), device=DEV)).shape == (0, 7, 16, H, W)
    img1 = torch.zeros((2, 6, H, W), device=DEV)
    flow1 = torch.zeros((2, 4, H, W), device=DEV)
    assert ssm_b200.fuse_from_flow(img1, flow1, torch.zeros((2, 0, 5, H, W), device=DEV), torch.zeros((2, 0), device=DEV)).shape == (2, 0, 3, H, W)
    # a caller-owned result buffer comes back as it is
    buf = torch.empty((0, 1, 16, H, W), device=DEV)
    assert ssm_b200.flow_pack(img.detach(), flow.detach(), t, out=buf) is buf
