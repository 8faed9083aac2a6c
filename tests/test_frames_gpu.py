"""GPU suite for the frame pre/post kernels: ssm_frames_from_u8 /
ssm_frames_to_u8 against the restated reference steps (oracle/torch_oracle.py), bit-exact."""
import numpy as np
import pytest
import torch

import ssm_b200
from oracle import torch_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _images(T, H, W, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, size=(T, H, W, 3)).astype(np.uint8)
    img[0, 0, 0] = (0, 128, 255)      # extremes
    return img


@pytest.mark.parametrize("H,W", [(720, 1280), (1080, 1920), (100, 52), (64, 96)])
@pytest.mark.parametrize("lut_device", ["cpu", DEV])
def test_from_u8_matches_visualize_path(H, W, lut_device):
    """uint8 BGR -> padded, normalised planar frames + RGBx, vs load_batch + normalize_tensor run with
    torch on the same device as the table (CPU: true division; CUDA: reciprocal multiply)."""
    bgr = _images(2, H, W, seed=H + W)
    want = torch_oracle.load_batch_and_normalize(bgr, device=lut_device)[0].cpu()       # T x 3 x H32 x W32
    lut = ssm_b200.normalisation_lut(device=lut_device)
    planar, rgbx, (top, left) = ssm_b200.frames_from_u8(torch.from_numpy(bgr).to(DEV), order="bgr", pad_mode="before",
                                                        lut=lut, want_rgbx=True)
    assert planar.shape == want.shape and (top, left) == ((want.shape[2] - H) // 2, (want.shape[3] - W) // 2)
    assert torch.equal(planar.cpu(), want), "max err %.3e" % (planar.cpu() - want).abs().max().item()
    assert torch.equal(rgbx[..., :3].permute(0, 3, 1, 2).cpu(), want)
    assert rgbx[..., 3].abs().max().item() == 0
    # the RGBx output of a frame pair is what ssm_pack_frames makes from the planar pair
    pair = planar.reshape(1, 6, *planar.shape[-2:])
    assert torch.equal(ssm_b200.pack_frames(pair), rgbx.view(1, 2, *rgbx.shape[1:]))


def test_from_u8_matches_reader_path():
    """data-loader flavour: float64 normalisation, zero padding AFTER normalising (720 -> 736)."""
    rgb = _images(3, 720, 1280, seed=5)
    want = torch_oracle.reader_normalize_and_pad(rgb, 8)
    lut = ssm_b200.normalisation_lut(style="reader", device=DEV)
    planar, _, (top, left) = ssm_b200.frames_from_u8(torch.from_numpy(rgb).to(DEV), order="rgb", pad_mode="after", lut=lut)
    assert (top, left) == (8, 0)
    assert torch.equal(planar.cpu(), want)


def test_from_u8_bf16_storage():
    bgr = _images(2, 64, 96, seed=9)
    want = torch_oracle.load_batch_and_normalize(bgr)[0]
    planar, rgbx, _ = ssm_b200.frames_from_u8(torch.from_numpy(bgr).to(DEV), lut=ssm_b200.normalisation_lut(device="cpu"),
                                              dtype=torch.bfloat16, want_rgbx=True)
    assert torch.equal(planar.cpu(), want.to(torch.bfloat16))
    assert torch.equal(rgbx[..., :3].permute(0, 3, 1, 2).cpu(), want.to(torch.bfloat16))


@pytest.mark.parametrize("saturate", [False, True])
def test_to_u8_matches_eval_path(saturate):
    """crop + de-normalise + astype(uint8): in-range values bit-exact; out-of-range values wrap like
    numpy's conversion (saturate=False) or clamp (saturate=True)."""
    H, W, h_in, w_in = 736, 1280, 720, 1280
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, H, W, generator=g) * 1.3          # a good share falls outside [0, 255] after de-normalising
    want = torch_oracle.crop_denormalize_u8(x, 8, 0, h_in, w_in)
    got = ssm_b200.frames_to_u8(x.to(DEV), top=8, left=0, h_out=h_in, w_out=w_in, saturate=saturate).cpu().numpy()
    if not saturate:
        assert np.array_equal(got, want)
    else:
        v = ((x.permute(0, 2, 3, 1)[:, 8:8 + h_in] * torch.tensor(torch_oracle.PIXEL_STD) + torch.tensor(torch_oracle.PIXEL_MEAN)) * 255.0)
        inside = ((v >= 0) & (v < 256)).numpy()
        assert np.array_equal(got[inside], want[inside])
        assert (got[(v < 0).numpy()] == 0).all() and (got[(v >= 256).numpy()] == 255).all()
    bgr = ssm_b200.frames_to_u8(x.to(DEV), top=8, left=0, h_out=h_in, w_out=w_in, order="bgr", saturate=saturate).cpu().numpy()
    assert np.array_equal(bgr[..., ::-1], got)


def test_round_trip_u8():
    """size-independent property at full size: from_u8 -> to_u8 reproduces every byte (1080p)."""
    bgr = _images(2, 1080, 1920, seed=11)
    planar, _, (top, left) = ssm_b200.frames_from_u8(torch.from_numpy(bgr).to(DEV), order="bgr")
    back = ssm_b200.frames_to_u8(planar, top=top, left=left, h_out=1080, w_out=1920, order="bgr", saturate=True)
    diff = (back.cpu().numpy().astype(np.int16) - bgr.astype(np.int16))
    # (v/255 - m)/s*s + m)*255 truncates: off by one below for the values whose round trip lands just under v
    assert diff.max() <= 0 and diff.min() >= -1


# ---- against outputs of the reference itself (tests/golden/frames_prepost.npz, make_golden_frames.py) -------
def _prepost():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_prepost.npz"))


def test_from_u8_matches_reference_fixture():
    """Interpolator.load_batch + normalize_tensor (run unmodified on CPU) on a ragged 45 x 70 clip."""
    d = _prepost()
    want = torch.from_numpy(d["vis_normalised"])[0]                      # T x 3 x 64 x 96
    planar, rgbx, (top, left) = ssm_b200.frames_from_u8(torch.from_numpy(d["vis_bgr_u8"]).to(DEV), order="bgr",
                                                        pad_mode="before", lut=ssm_b200.normalisation_lut(device="cpu"),
                                                        want_rgbx=True)
    assert (top, left) == (9, 13)
    assert torch.equal(planar.cpu(), want)
    assert torch.equal(rgbx[..., :3].permute(0, 3, 1, 2).cpu(), want)


def test_from_u8_matches_reference_reader_fixture():
    """augmentations.Normalize + ToTensor + EvalPad (run unmodified on CPU)."""
    d = _prepost()
    want = torch.from_numpy(d["reader_out"])
    pad = int(d["reader_pad"][0])
    rgb = torch.from_numpy(d["reader_rgb_u8"]).to(DEV)
    # the reader pads rows only (ZeroPad2d([0, 0, pad, pad])): 40 + 2 * 12 = 64 rows, 64 columns
    planar, _, (top, left) = ssm_b200.frames_from_u8(rgb, order="rgb", pad_mode="after",
                                                     lut=ssm_b200.normalisation_lut(style="reader", device=DEV))
    assert (top, left) == (pad, 0) and planar.shape == want.shape
    assert torch.equal(planar.cpu(), want)


def test_to_u8_matches_reference_fixture():
    """Evaluator.convert_tensor_to_numpy_image (get_crop + denormalize + astype(uint8), run unmodified on
    CPU): values that de-normalise into [0, 256) must be bit-exact; outside that range astype(uint8) of a
    float is implementation-defined in numpy, the kernel wraps modulo 256 like x86 numpy does."""
    d = _prepost()
    h0, w0, h, w = [int(v) for v in d["eval_crop"]]
    x = torch.from_numpy(d["eval_in"])
    got = ssm_b200.frames_to_u8(x.to(DEV), top=h0, left=w0, h_out=h, w_out=w).cpu().numpy()
    assert got.shape == d["eval_u8"].shape
    assert np.array_equal(got, d["eval_u8"])
    # visualiser flavour: denormalize_tensor then astype(uint8) on the full frame (visualize_interpolation.py:225-232, 264-268)
    xin = torch.from_numpy(d["vis_denorm_in"])[0]
    want = d["vis_denorm_out"][0].transpose(0, 2, 3, 1).astype(np.uint8)
    got = ssm_b200.frames_to_u8(xin.to(DEV)).cpu().numpy()
    v = d["vis_denorm_out"][0].transpose(0, 2, 3, 1)
    inside = (v >= 0) & (v < 256)
    assert np.array_equal(got[inside], want[inside])
