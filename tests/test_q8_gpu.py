"""GPU suite (-m gpu): the 8-bit-frame entry points (ssm_quads_from_u8, ssm_flow_pack_fwd_q8[_nhwc],
ssm_fuse_flow_fwd_q8[_u8], ssm_synthesize_host_u8) through the C ABI against
  * fixtures produced by the reference's own image loading + normalisation + path (tests/golden/make_golden_q8.py),
  * the C oracle run on the normalised frames (the reference order: normalise, then interpolate),
  * the fp32 kernels of this library on the same frames,
and full-size properties at BASELINE.json's size.  Tolerance: 1e-5 (north_star, fp32); the pass-through channels and
the estimated flows are bit-identical.
"""
import numpy as np
import pytest
import torch

import ssm_b200
from oracle import c_oracle, torch_oracle
from ssm_b200 import q8, synthetic
from util import assert_close_fp32, load_golden, max_err, q8_cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODES = [("cpu", c_oracle.COORD_DIV), ("cuda", c_oracle.COORD_RCP)]


def _prepared(bgr_u8):
    """uint8 BGR images (numpy F x h x w x 3) -> device tensors (planar, quads, norm, (top, left))."""
    images = torch.from_numpy(np.ascontiguousarray(bgr_u8)).to(DEV)
    # the fixtures come from a CPU run of the reference: the table is filled with the CPU bit pattern of its expression
    return q8.prepare(images, order="bgr", lut=ssm_b200.normalisation_lut(device="cpu"))


@pytest.mark.parametrize("name", q8_cases())
def test_reference_fixture(name):
    d = load_golden(name)
    B, N = d["flow4"].shape[0], d["t"].numel()
    planar, quads, norm, _ = _prepared(d["bgr_u8"].numpy())
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W)
    assert torch.equal(img6.cpu(), d["img6"]), "normalised frames are not bit-identical to the reference's"
    t = d["t"].view(1, N).expand(B, N).contiguous()
    flow4, out5 = d["flow4"].to(DEV), d["out5"].to(DEV)
    in16 = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N)
    assert torch.equal(in16[:, :, 6:10].cpu(), d["in16"][:, :, 6:10]), "estimated flows not bit-identical"
    assert torch.equal(in16[:, :, 0:3].cpu(), d["in16"][:, :, 0:3]) and torch.equal(in16[:, :, 13:16].cpu(), d["in16"][:, :, 13:16])
    assert_close_fp32(in16, d["in16"], "compute_inputs from uint8 frames")
    frames = q8.fuse_from_flow(quads, flow4, out5, t, norm)
    assert_close_fp32(frames, d["frames"], "compute_output_image from uint8 frames")
    # the table-only form (no planar frames read): the reference's own pass-through channels, bit for bit
    lut = ssm_b200.normalisation_lut(device="cpu").to(DEV)
    in16_t = q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N, lut=lut)
    assert torch.equal(in16_t, in16), "compute_inputs with the pass-through channels from the tables differs"


def _u8_images(F, h, w, seed, smooth):
    if smooth:
        x = synthetic.frames(F, h, w, n_frames=1, seed=seed, smooth=True)
        x = (x - x.amin()) / (x.amax() - x.amin())
        return (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (F, h, w, 3), dtype=torch.uint8, generator=g)


@pytest.mark.parametrize("mode_name,mode", MODES)
@pytest.mark.parametrize("B,N,h,w,kind,smooth,flow_px", [
    (2, 3, 64, 96, "smooth", True, 8.0),
    (1, 7, 45, 77, "noise", False, 8.0),       # ragged source, padded to 64 x 96
    (2, 2, 33, 30, "border", True, 8.0),       # samples cross every border of the padded frame
    (1, 2, 8, 1920, "integer", False, 8.0),
    (1, 1, 2, 2, "zero", False, 8.0),
    (2, 2, 352, 352, "smooth", True, 20.0),    # +-100 px flows
])
def test_q8_vs_c_oracle(mode_name, mode, B, N, h, w, kind, smooth, flow_px):
    images = _u8_images(2 * B, h, w, seed=300 + h + w, smooth=smooth)
    planar, quads, norm, (top, left) = q8.prepare(images.to(DEV), order="rgb", lut=ssm_b200.normalisation_lut(device="cpu"))
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=flow_px, seed=301 + h, kind=kind)
    out5 = synthetic.unet_out5(B, N, H, W, seed=302 + w)
    t = synthetic.timesteps(B, N)
    in16 = q8.flow_pack(img6, quads, flow4.to(DEV), t, norm, n_timesteps=N, coord_mode=mode_name)
    frames = q8.fuse_from_flow(quads, flow4.to(DEV), out5.to(DEV), t, norm, coord_mode=mode_name)
    img6_c = img6.cpu()
    worst16 = worst3 = 0.0
    for n in range(N):
        r16 = c_oracle.compute_inputs(img6_c, flow4, t[:, n], coord_mode=mode)
        assert torch.equal(in16[:, n, 6:10].cpu(), r16[:, 6:10]), "estimated flows not bit-identical"
        assert torch.equal(in16[:, n, 0:3].cpu(), r16[:, 0:3]) and torch.equal(in16[:, n, 13:16].cpu(), r16[:, 13:16])
        worst16 = max(worst16, assert_close_fp32(in16[:, n], r16, "q8 flow_pack n=%d" % n))
        r3 = c_oracle.compute_output_image(img6_c, r16, out5[:, n].contiguous(), t[:, n], coord_mode=mode)
        worst3 = max(worst3, assert_close_fp32(frames[:, n], r3, "q8 fuse n=%d" % n))
    print("q8 vs C oracle %s %dx%d: warped %.2e fused %.2e" % (mode_name, H, W, worst16, worst3))


@pytest.mark.parametrize("coord_mode", ["cpu", "cuda"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_q8_flow_pack_from_tables_is_bit_identical(dtype, coord_mode):
    """ssm_flow_pack_fwd_q8_lut: the pass-through channels looked up from the tables' own bytes instead of read from the
    planar frames -- every one of the 16 channels must be the same bits (random bytes, a ragged source padded to a
    multiple of 32, so the padding columns and rows are covered), in both storage types."""
    B, N, h, w = 2, 3, 45, 78
    images = _u8_images(2 * B, h, w, seed=77, smooth=False).to(DEV)
    lut = ssm_b200.normalisation_lut(device=DEV)
    planar, quads, norm, _ = q8.prepare(images, order="rgb", lut=lut)
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W).to(dtype)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=9.0, seed=78, device=DEV).to(dtype)
    t = synthetic.timesteps(B, N, device=DEV)
    a = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N, coord_mode=coord_mode)
    b = q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N, coord_mode=coord_mode, lut=lut)
    assert a.dtype == b.dtype == dtype and torch.equal(a, b)
    buf = torch.empty_like(a)
    assert q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N, coord_mode=coord_mode, lut=lut, out=buf) is buf
    assert torch.equal(buf, a)
    with pytest.raises(RuntimeError, match="lut"):
        q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N)
    with pytest.raises(RuntimeError, match="lut"):
        q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N, lut=lut[:, :128].contiguous())
    with pytest.raises(RuntimeError, match="planar"):
        q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=N, lut=lut, channels_last_dtype=torch.bfloat16)


def test_q8_layouts_and_bf16_unet_output():
    """channels-last / bf16 stage-2 input and a bf16 U-Net output, against the planar fp32 q8 kernels"""
    B, N, h, w = 2, 3, 40, 62
    images = _u8_images(2 * B, h, w, seed=77, smooth=True)
    planar, quads, norm, _ = q8.prepare(images.to(DEV))
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=8.0, seed=78).to(DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=79).to(DEV)
    t = synthetic.timesteps(B, N)
    ref16 = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N)
    cl32 = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N, channels_last_dtype=torch.float32)
    assert cl32.stride() == (N * 16 * H * W, 16 * H * W, 1, 16 * W, 16) and torch.equal(cl32, ref16)
    cl16 = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N, channels_last_dtype=torch.bfloat16)
    assert torch.equal(cl16, ref16.bfloat16()), "channels-last bf16 is not the fp32 result rounded once"
    y16 = out5.bfloat16()
    a = q8.fuse_from_flow(quads, flow4, y16, t, norm)
    b = q8.fuse_from_flow(quads, flow4, y16.float(), t, norm)
    assert torch.equal(a, b), "a bf16 U-Net output must read as its fp32 widening"


@pytest.mark.parametrize("order", ["bgr", "rgb"])
def test_q8_uint8_output_matches_unfused_pipeline(order):
    """ssm_fuse_flow_fwd_q8_u8 == ssm_fuse_flow_fwd_q8 followed by ssm_frames_to_u8 (crop, de-normalise, clamp)"""
    B, N, h, w = 2, 2, 45, 70
    images = _u8_images(2 * B, h, w, seed=91, smooth=True)
    planar, quads, norm, (top, left) = q8.prepare(images.to(DEV), order=order)
    H, W = planar.shape[-2:]
    flow4 = synthetic.flows(B, H, W, 4, flow_px=6.0, seed=92).to(DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=93).to(DEV)
    t = synthetic.timesteps(B, N)
    frames = q8.fuse_from_flow(quads, flow4, out5, t, norm)
    want = ssm_b200.frames_to_u8(frames.view(B * N, 3, H, W), top=top, left=left, h_out=h, w_out=w, order=order, saturate=True)
    got = q8.fuse_from_flow_to_u8(quads, flow4, out5, t, norm, crop=(top, left, h, w), order=order, saturate=True)
    assert torch.equal(got.view(B * N, h, w, 3), want)
    # zero flow + a visibility that selects frame 0 reproduces the source image exactly
    zf = torch.zeros_like(flow4)
    y0 = torch.zeros_like(out5)
    y0[:, :, 0] = -40.0                                      # V_t<-1 = sigmoid(-40) = 0: only frame 0 contributes
    back = q8.fuse_from_flow_to_u8(quads, zf, y0, t, norm, crop=(top, left, h, w), order=order, saturate=True)
    src0 = images.view(B, 2, h, w, 3)[:, 0].to(DEV)
    diff = (back.int() - src0.unsqueeze(1).int()).abs().max().item()
    assert diff <= 1, "zero-flow round trip moved a byte by %d" % diff      # (x*std+mean)*255 truncates: off by one at most


def test_q8_host_entry_matches_device_path():
    B, N, h, w = 3, 3, 45, 70
    images = _u8_images(2 * B, h, w, seed=55, smooth=True).view(B, 2, h, w, 3).contiguous()
    planar, quads, norm, (top, left) = q8.prepare(images.view(2 * B, h, w, 3).to(DEV))
    H, W = planar.shape[-2:]
    flow4 = synthetic.flows(B, H, W, 4, flow_px=6.0, seed=56)
    out5 = synthetic.unet_out5(B, N, H, W, seed=57)
    t = synthetic.timesteps(B, N)
    want = q8.fuse_from_flow_to_u8(quads, flow4.to(DEV), out5.to(DEV), t, norm, crop=(top, left, h, w))
    got = q8.synthesize_host(images.pin_memory(), flow4.pin_memory(), out5.pin_memory(), t)
    assert torch.equal(got, want.cpu())
    got16 = q8.synthesize_host(images.pin_memory(), flow4.pin_memory(), out5.bfloat16().pin_memory(), t)
    want16 = q8.fuse_from_flow_to_u8(quads, flow4.to(DEV), out5.bfloat16().to(DEV), t, norm, crop=(top, left, h, w))
    assert torch.equal(got16, want16.cpu())


def test_q8_full_size_against_fp32_kernels():
    """BASELINE.json configs[1] size (4 of the 16 pairs): the 8-bit path against this library's fp32 kernels on the
    normalised frames -- two independent gather implementations -- plus one C-oracle slice."""
    B, N, h, w = 4, 7, 1080, 1920
    images = _u8_images(2 * B, h, w, seed=11, smooth=True)
    planar, quads, norm, _ = q8.prepare(images.to(DEV))
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=12, device=DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=13, device=DEV)
    t = synthetic.timesteps(B, N, device=DEV)
    in16 = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N)
    frames = q8.fuse_from_flow(quads, flow4, out5, t, norm)
    ref16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N)
    ref3 = ssm_b200.fuse_from_flow(img6, flow4, out5, t)
    assert torch.equal(in16[:, :, 6:10], ref16[:, :, 6:10]) and torch.equal(in16[:, :, 0:3], ref16[:, :, 0:3])
    e16, e3 = max_err(in16, ref16), max_err(frames, ref3)
    print("q8 vs fp32 kernels at 1088x1920: warped %.2e fused %.2e" % (e16, e3))
    assert e16 <= 1e-5 and e3 <= 1e-5
    # one oracle slice: rows 500..507 of pair 1, timestep 3 (the oracle computes whole frames: restrict the comparison)
    r16 = c_oracle.compute_inputs(img6[1:2].cpu(), flow4[1:2].cpu(), t[1:2, 3].cpu())
    assert_close_fp32(in16[1, 3, :, 500:508], r16[0, :, 500:508], "q8 flow_pack vs C oracle at full size")


def test_q8_bf16_storage():
    """bf16 storage of flows / frames / stage-2 input / result with fp32 arithmetic, against the C oracle on the same
    bf16-rounded inputs (north_star: <= 2e-2, + 2^-7 relative above magnitude 1: util.assert_close_bf16)."""
    from util import assert_close_bf16
    B, N, h, w = 2, 3, 64, 96
    images = _u8_images(2 * B, h, w, seed=21, smooth=True)
    planar, quads, norm, _ = q8.prepare(images.to(DEV), order="rgb", lut=ssm_b200.normalisation_lut(device="cpu"))
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=8.0, seed=22).bfloat16()
    out5 = synthetic.unet_out5(B, N, H, W, seed=23).bfloat16()
    t = synthetic.timesteps(B, N)
    in16 = q8.flow_pack(img6.bfloat16(), quads, flow4.to(DEV), t, norm, n_timesteps=N)
    frames = q8.fuse_from_flow(quads, flow4.to(DEV), out5.to(DEV), t, norm)
    assert in16.dtype == torch.bfloat16 and frames.dtype == torch.bfloat16
    img6_c = img6.cpu()               # the gathers read the exact bytes: the oracle warps the unrounded frames
    for n in range(N):
        r16 = c_oracle.compute_inputs(img6_c, flow4.float(), t[:, n])
        # the estimated flows are rounded to bf16 before they are used: warp the oracle with the rounded values
        e = r16[:, 6:10].bfloat16().float()
        assert torch.equal(in16[:, n, 6:10].float().cpu(), e), "estimated flows are not the bf16 rounding of the oracle's"
        w1 = c_oracle.warp(img6_c[:, 3:6].contiguous(), e[:, 0:2].contiguous())
        w0 = c_oracle.warp(img6_c[:, 0:3].contiguous(), e[:, 2:4].contiguous())
        assert_close_bf16(in16[:, n, 3:6], w1, "q8 bf16 warped I1 n=%d" % n)
        assert_close_bf16(in16[:, n, 10:13], w0, "q8 bf16 warped I0 n=%d" % n)
        assert_close_bf16(in16[:, n, 0:3], img6_c[:, 3:6], "q8 bf16 pass-through n=%d" % n)
        x16 = r16.clone()
        x16[:, 6:10] = e
        r3 = c_oracle.compute_output_image(img6_c, x16, out5[:, n].float().contiguous(), t[:, n])
        assert_close_bf16(frames[:, n], r3, "q8 bf16 fused n=%d" % n)


def test_q8_4k_31_timesteps_against_fp32_kernels():
    """BASELINE.json configs[4] size: one 2160x3840 pair (padded to 2176x3840), t = k/32 -- the timesteps one rank of
    eight gets (4) plus the last one -- against the fp32 kernels, and the oracle on a band of rows."""
    h, w = 2160, 3840
    images = _u8_images(2, h, w, seed=21, smooth=True)
    planar, quads, norm, _ = q8.prepare(images.to(DEV))
    H, W = planar.shape[-2:]
    assert (H, W) == (2176, 3840)
    img6 = planar.view(1, 6, H, W)
    flow4 = synthetic.flows(1, H, W, 4, flow_px=20.0, seed=22, device=DEV)
    t = torch.tensor([[1 / 32, 2 / 32, 3 / 32, 4 / 32, 31 / 32]], device=DEV)
    N = t.shape[1]
    out5 = synthetic.unet_out5(1, N, H, W, seed=23, device=DEV)
    in16 = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N)
    frames = q8.fuse_from_flow(quads, flow4, out5, t, norm)
    ref16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N)
    ref3 = ssm_b200.fuse_from_flow(img6, flow4, out5, t)
    assert torch.equal(in16[:, :, 6:10], ref16[:, :, 6:10])
    assert max_err(in16, ref16) <= 1e-5 and max_err(frames, ref3) <= 1e-5
    assert torch.isfinite(frames).all()
    r16 = c_oracle.compute_inputs(img6.cpu(), flow4.cpu(), t[:, 4].cpu())
    assert_close_fp32(in16[0, 4, :, 2000:2008], r16[0, :, 2000:2008], "q8 flow_pack vs C oracle at 4K")


def test_q8_non_finite_flows_and_argument_errors():
    """Samples with non-finite or absurd coordinates read nothing and give zeros (the fp32 kernels and, in practice, the
    reference do the same); wrong arguments are refused loudly -- there is no fallback path."""
    B, N, h, w = 1, 2, 40, 72
    images = _u8_images(2 * B, h, w, seed=31, smooth=False)
    planar, quads, norm, _ = q8.prepare(images.to(DEV))
    H, W = planar.shape[-2:]
    img6 = planar.view(B, 6, H, W)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=4.0, seed=32, device=DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=33, device=DEV)
    t = synthetic.timesteps(B, N, device=DEV)
    bad = flow4.clone()
    bad[0, :, 5, 7] = float("inf")
    bad[0, :, 6, 8] = float("nan")
    bad[0, :, 7, 9] = 3.0e30
    bad[0, :, 8, 10] = -1.0e9
    in16 = q8.flow_pack(img6, quads, bad, t, norm, n_timesteps=N)
    assert torch.isfinite(in16[:, :, 3:6]).all() and torch.isfinite(in16[:, :, 10:13]).all()
    for (y, x) in ((5, 7), (6, 8), (7, 9), (8, 10)):
        assert in16[0, :, 3:6, y, x].abs().max() == 0 and in16[0, :, 10:13, y, x].abs().max() == 0
    good = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=N)
    mask = torch.ones((H, W), dtype=torch.bool, device=DEV)
    for (y, x) in ((5, 7), (6, 8), (7, 9), (8, 10)):
        mask[y, x] = False
    assert torch.equal(in16[:, :, 3:6][..., mask], good[:, :, 3:6][..., mask])          # every other pixel untouched
    frames = q8.fuse_from_flow(quads, bad, out5, t, norm)
    assert torch.isfinite(frames).all()
    with pytest.raises(RuntimeError):                     # tables of another size
        q8.flow_pack(img6, quads[:, :-1].contiguous(), flow4, t, norm, n_timesteps=N)
    with pytest.raises(RuntimeError):                     # CPU tensors: no fallback
        q8.flow_pack(img6.cpu(), quads, flow4.cpu(), t.cpu(), norm, n_timesteps=N)
    with pytest.raises(RuntimeError):                     # inference only
        q8.flow_pack(img6, quads, flow4.clone().requires_grad_(True), t, norm, n_timesteps=N)
    with pytest.raises(RuntimeError):                     # t count (a single value would be broadcast; three for two is an error)
        q8.fuse_from_flow(quads, flow4, out5, torch.tensor([[0.25, 0.5, 0.75]], device=DEV), norm)
    with pytest.raises(AssertionError):                   # validators.py:9-11 on a host-side t
        q8.flow_pack(img6, quads, flow4, torch.tensor([[0.5, 1.0]]), norm, n_timesteps=N)


def test_q8_and_fp32_kernels_on_random_shapes_vs_c_oracle():
    """Twelve seeded random problems (image sizes 2..140 that are not multiples of anything, 1..5 timesteps, flows from
    sub-pixel to several image widths, both coordinate modes): the 8-bit-frame kernels and the fp32 kernels against the C
    oracle on the normalised frames, 1e-5; estimated flows bit-identical."""
    rng = np.random.default_rng(2024)
    for case in range(12):
        B, N = int(rng.integers(1, 3)), int(rng.integers(1, 6))
        h, w = int(rng.integers(2, 141)), int(rng.integers(2, 141))
        flow_px = float(rng.choice([0.3, 2.0, 9.0, 40.0, 300.0]))
        kind = str(rng.choice(["smooth", "noise", "border"]))
        mode_name, mode = MODES[case % 2]
        images = _u8_images(2 * B, h, w, seed=900 + case, smooth=bool(case % 3))
        planar, quads, norm, _ = q8.prepare(images.to(DEV), order="rgb", lut=ssm_b200.normalisation_lut(device="cpu"))
        H, W = planar.shape[-2:]
        img6 = planar.view(B, 6, H, W)
        flow4 = synthetic.flows(B, H, W, 4, flow_px=flow_px, seed=901 + case, kind=kind)
        out5 = synthetic.unet_out5(B, N, H, W, seed=902 + case)
        t = synthetic.timesteps(B, N)
        in16_q = q8.flow_pack(img6, quads, flow4.to(DEV), t, norm, n_timesteps=N, coord_mode=mode_name)
        fr_q = q8.fuse_from_flow(quads, flow4.to(DEV), out5.to(DEV), t, norm, coord_mode=mode_name)
        in16_f = ssm_b200.flow_pack(img6, flow4.to(DEV), t.to(DEV), n_timesteps=N, coord_mode=mode_name)
        fr_f = ssm_b200.fuse_from_flow(img6, flow4.to(DEV), out5.to(DEV), t.to(DEV), coord_mode=mode_name)
        img6_c = img6.cpu()
        what = "case %d (B=%d N=%d %dx%d -> %dx%d, %s flow x %g, %s)" % (case, B, N, h, w, H, W, kind, flow_px, mode_name)
        for n in range(N):
            r16 = c_oracle.compute_inputs(img6_c, flow4, t[:, n], coord_mode=mode)
            r3 = c_oracle.compute_output_image(img6_c, r16, out5[:, n].contiguous(), t[:, n], coord_mode=mode)
            assert torch.equal(in16_q[:, n, 6:10].cpu(), r16[:, 6:10]) and torch.equal(in16_f[:, n, 6:10].cpu(), r16[:, 6:10]), what
            assert_close_fp32(in16_q[:, n], r16, "q8 compute_inputs, " + what)
            assert_close_fp32(in16_f[:, n], r16, "fp32 compute_inputs, " + what)
            assert_close_fp32(fr_q[:, n], r3, "q8 compute_output_image, " + what)
            assert_close_fp32(fr_f[:, n], r3, "fp32 compute_output_image, " + what)
