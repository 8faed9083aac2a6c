"""GPU suite: the layouts either side of the stage-2 U-Net.
  * ssm_flow_pack_fwd_nhwc: compute_inputs written channels-last, in fp32 or bf16 -- must be the planar
    result (itself parity-tested against the reference) rounded once to the output dtype, bit for bit;
  * ssm_fuse_flow_fwd_mixed: compute_output_image reading a bf16 U-Net output next to fp32 frames -- must be
    the fp32 call on out5.float(), bit for bit;
  * FullModel.interpolate with a channels-last U-Net under bf16 autocast uses both.
"""
import pytest
import torch

import ssm_b200
from ssm_b200 import functional as F_ssm
from ssm_b200 import synthetic
from ssm_b200.superslomo_r import FullModel
from util import seeded_unets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B, N, H, W, seed):
    img6 = synthetic.frames(B, H, W, seed=seed).to(DEV)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=4.0, seed=seed + 1).to(DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=seed + 2).to(DEV)
    t = synthetic.timesteps(B, N).to(DEV)
    return img6, flow4, out5, t


@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("B,N,H,W,packed", [(2, 3, 64, 96, True), (1, 1, 37, 70, False), (2, 7, 40, 100, True)])
@pytest.mark.parametrize("mode", ["cpu", "cuda"])
def test_channels_last_compute_inputs_is_the_planar_result(in_dtype, out_dtype, B, N, H, W, packed, mode):
    img6, flow4, _, t = _inputs(B, N, H, W, seed=900 + N)
    img6, flow4 = img6.to(in_dtype), flow4.to(in_dtype)
    rgbx = ssm_b200.pack_frames(img6) if packed else None
    planar = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, coord_mode=mode, packed=rgbx)
    if not packed and N >= 2:
        rgbx = None
    nhwc = ssm_b200.flow_pack_channels_last(img6, flow4, t, n_timesteps=N, dtype=out_dtype, coord_mode=mode, packed=rgbx)
    assert nhwc.shape == (B, N, 16, H, W) and nhwc.dtype == out_dtype
    assert nhwc.stride() == (N * 16 * H * W, 16 * H * W, 1, 16 * W, 16)
    assert nhwc.view(B * N, 16, H, W).is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(nhwc, planar.to(out_dtype)), "channels-last compute_inputs differs from the planar result"


def test_channels_last_compute_inputs_out_buffer_and_errors():
    B, N, H, W = 2, 2, 32, 64
    img6, flow4, _, t = _inputs(B, N, H, W, seed=950)
    buf = torch.empty((B, N, H, W, 16), dtype=torch.bfloat16, device=DEV).permute(0, 1, 4, 2, 3)
    got = ssm_b200.flow_pack_channels_last(img6, flow4, t, n_timesteps=N, dtype=torch.bfloat16, out=buf)
    assert got.data_ptr() == buf.data_ptr()
    assert torch.equal(got, ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N).bfloat16())
    with pytest.raises(RuntimeError):       # a planar buffer is not the promised layout
        ssm_b200.flow_pack_channels_last(img6, flow4, t, n_timesteps=N, dtype=torch.bfloat16,
                                         out=torch.empty((B, N, 16, H, W), dtype=torch.bfloat16, device=DEV))
    with pytest.raises(RuntimeError):       # inference only
        ssm_b200.flow_pack_channels_last(img6, flow4.clone().requires_grad_(True), t, n_timesteps=N)
    with pytest.raises(TypeError):          # bf16 frames cannot be widened
        ssm_b200.flow_pack_channels_last(img6.bfloat16(), flow4.bfloat16(), t, n_timesteps=N, dtype=torch.float32)
    with pytest.raises(RuntimeError):       # no CPU fallback
        ssm_b200.flow_pack_channels_last(img6.cpu(), flow4.cpu(), t.cpu(), n_timesteps=N)


@pytest.mark.parametrize("B,N,H,W", [(2, 3, 64, 96), (1, 7, 37, 70)])
@pytest.mark.parametrize("mode", ["cpu", "cuda"])
def test_bf16_unet_output_next_to_fp32_frames(B, N, H, W, mode):
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=970)
    y = out5.bfloat16()
    want = ssm_b200.fuse_from_flow(img6, flow4, y.float(), t, coord_mode=mode)
    got = ssm_b200.fuse_from_flow(img6, flow4, y, t, coord_mode=mode)
    assert got.dtype == torch.float32
    assert torch.equal(got, want)
    # a channel-sliced / strided bf16 view is accepted like everywhere else
    wide = torch.zeros((B, N, 8, H, W), dtype=torch.bfloat16, device=DEV)
    wide[:, :, 2:7] = y
    assert torch.equal(ssm_b200.fuse_from_flow(img6, flow4, wide[:, :, 2:7], t, coord_mode=mode), want)
    with pytest.raises(RuntimeError):       # inference only
        ssm_b200.fuse_from_flow(img6, flow4, y.clone().requires_grad_(True), t, coord_mode=mode)


def test_interpolate_uses_the_unet_layouts(monkeypatch):
    """channels-last U-Nets under bf16 autocast: compute_inputs hands conv1a a channels-last bf16 tensor (no
    conversion pass) and compute_output_image reads the bf16 output; same frames as the generic plumbing
    up to the U-Nets' own bf16 noise."""
    B, H, W, n_t = 2, 64, 96, 3
    s1, s2 = seeded_unets(123, DEV)
    m = FullModel(cfg=None, stage1_model=s1, stage2_model=s2).eval()
    m.stage1_model.set_channels_last()
    m.stage2_model.set_channels_last()
    frames = synthetic.frames(B, H, W, n_frames=2, seed=31).view(B, 2, 3, H, W).to(DEV)
    tv = torch.arange(1, n_t + 1, dtype=torch.float32) / (n_t + 1)
    seen = {"nhwc": 0, "mixed": 0, "conv_in": []}
    real_nhwc, real_mixed = F_ssm.flow_pack_channels_last, F_ssm._fuse_from_flow_mixed

    def spy_nhwc(*a, **k):
        seen["nhwc"] += 1
        return real_nhwc(*a, **k)

    def spy_mixed(*a, **k):
        seen["mixed"] += 1
        return real_mixed(*a, **k)

    monkeypatch.setattr(F_ssm, "flow_pack_channels_last", spy_nhwc)
    monkeypatch.setattr(F_ssm, "_fuse_from_flow_mixed", spy_mixed)
    real_block = m.stage2_model._block

    def spy_block(seq, x, **kw):               # what the first convolution block of stage 2 is handed
        if seq is m.stage2_model.conv1a:
            seen["conv_in"].append((x.dtype, x.is_contiguous(memory_format=torch.channels_last)))
        return real_block(seq, x, **kw)

    monkeypatch.setattr(m.stage2_model, "_block", spy_block)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fast = m.interpolate(frames, tv)
    monkeypatch.setattr(m.stage2_model, "_block", real_block)
    assert seen["nhwc"] == 1 and seen["mixed"] == 1
    assert seen["conv_in"] == [(torch.bfloat16, True)]
    # generic plumbing: planar fp32 compute_inputs, converted by the U-Net wrapper and autocast
    m.unet_layouts = False
    with torch.autocast("cuda", dtype=torch.bfloat16):
        slow = m.interpolate(frames, tv)
    assert fast.shape == slow.shape == (B, n_t, 3, H, W) and fast.dtype == torch.float32
    assert (fast - slow).abs().max().item() <= 5e-2
    # without autocast the channels-last tensor is fp32 and the result is the generic one
    m.unet_layouts = True
    a = m.interpolate(frames, tv)
    m.unet_layouts = False
    b = m.interpolate(frames, tv)
    assert (a - b).abs().max().item() <= 1e-3


@pytest.mark.parametrize("mode_name", ["cpu", "cuda"])
def test_layout_entry_points_against_the_c_oracle(mode_name):
    """The channels-last compute_inputs and the mixed-dtype compute_output_image DIRECTLY against the C restatement of
    the reference (oracle/ssm_oracle.c), not only against this library's planar kernels: fp32 channels-last within
    1e-5 with bit-identical estimated flows, bf16 channels-last = the oracle rounded once to bf16 up to one bf16 ulp,
    bf16 U-Net output = the oracle run on the widened values."""
    from oracle import c_oracle
    from util import assert_close_bf16, assert_close_fp32
    mode = {"cpu": c_oracle.COORD_DIV, "cuda": c_oracle.COORD_RCP}[mode_name]
    B, N, H, W = 2, 3, 40, 72
    img6, flow4, out5, t = _inputs(B, N, H, W, seed=990)
    nhwc32 = ssm_b200.flow_pack_channels_last(img6, flow4, t, n_timesteps=N, dtype=torch.float32, coord_mode=mode_name)
    nhwc16 = ssm_b200.flow_pack_channels_last(img6, flow4, t, n_timesteps=N, dtype=torch.bfloat16, coord_mode=mode_name)
    y16 = out5.bfloat16()
    mixed = ssm_b200.fuse_from_flow(img6, flow4, y16, t, coord_mode=mode_name)
    for n in range(N):
        r16 = c_oracle.compute_inputs(img6.cpu(), flow4.cpu(), t[:, n].cpu(), coord_mode=mode)
        assert torch.equal(nhwc32[:, n, 6:10].cpu(), r16[:, 6:10]), "estimated flows not bit-identical"
        assert_close_fp32(nhwc32[:, n], r16, "channels-last compute_inputs vs C oracle")
        assert_close_bf16(nhwc16[:, n], r16, "channels-last bf16 compute_inputs vs C oracle")
        r3 = c_oracle.compute_output_image(img6.cpu(), r16, y16[:, n].float().cpu().contiguous(), t[:, n].cpu(), coord_mode=mode)
        assert_close_fp32(mixed[:, n], r3, "compute_output_image with a bf16 U-Net output vs C oracle")
