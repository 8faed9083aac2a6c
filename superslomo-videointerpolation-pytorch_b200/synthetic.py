"""Seeded synthetic inputs for the synthesis path: there is no network for
datasets or checkpoints, so tests and bench.py drive the path with these.

  frames   uniform[0,1) RGB, low-passed (5x5 box, 4 passes), min-max rescaled, ImageNet-normalised
           with the reference's mean/std (configs/superslomo_original.ini:57-58)
  flows    randn at 1/8 resolution x `flow_px`, bilinearly upsampled (|flow| up to ~5*flow_px)
  out5     surrogate of the stage-2 U-Net output: channel 0 ~ N(0, 2^2) visibility logits,
           channels 1-4 ~ N(0, 0.5^2) px residual flows
  t        k/8 for k = 1..7 (evaluate_interpolation_results.py:204-211); k/32 for 31 timesteps
"""
import torch
import torch.nn.functional as F

PIXEL_MEAN = (0.485, 0.456, 0.406)
PIXEL_STD = (0.229, 0.224, 0.225)
SEED = 42  # configs/superslomo_original.ini:118-119


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def frames(B, H, W, n_frames=2, seed=SEED, device="cpu", smooth=True):
    """B x (3*n_frames) x H x W normalised frames (n_frames=2 gives the reference's 6-channel pair)."""
    g = _gen(seed, device)
    x = torch.rand((B * n_frames, 3, H, W), generator=g, device=device)
    if smooth:
        k = torch.full((3, 1, 5, 5), 1.0 / 25.0, device=device)
        for _ in range(4):
            x = F.conv2d(F.pad(x, (2, 2, 2, 2), mode="replicate"), k, groups=3)
        lo = x.amin(dim=(2, 3), keepdim=True)
        hi = x.amax(dim=(2, 3), keepdim=True)
        x = (x - lo) / (hi - lo + 1e-12)
    mean = torch.tensor(PIXEL_MEAN, device=device).view(1, 3, 1, 1)
    std = torch.tensor(PIXEL_STD, device=device).view(1, 3, 1, 1)
    x = (x - mean) / std
    return x.view(B, 3 * n_frames, H, W).contiguous()


def flows(B, H, W, channels=4, flow_px=20.0, seed=SEED + 1, device="cpu", kind="smooth"):
    """B x channels x H x W flow field.  kind: smooth | zero | integer | noise | border."""
    g = _gen(seed, device)
    if kind == "zero":
        return torch.zeros((B, channels, H, W), device=device)
    if kind == "integer":
        f = torch.empty((B, channels, H, W), device=device)
        f[:, 0::2] = 5.0
        f[:, 1::2] = -3.0
        return f
    if kind == "noise":
        return torch.randn((B, channels, H, W), generator=g, device=device)
    if kind == "border":
        # pushes samples across every image border to exercise the zero fill
        f = torch.randn((B, channels, H, W), generator=g, device=device) * 2.0
        f[:, 0::2] += torch.linspace(-1.5 * W, 1.5 * W, W, device=device).view(1, 1, 1, W) * 0.25
        f[:, 1::2] += torch.linspace(-1.5 * H, 1.5 * H, H, device=device).view(1, 1, H, 1) * 0.25
        return f.contiguous()
    h8, w8 = max(H // 8, 2), max(W // 8, 2)
    c = torch.randn((B, channels, h8, w8), generator=g, device=device) * flow_px
    return F.interpolate(c, size=(H, W), mode="bilinear", align_corners=False).contiguous()


def unet_out5(B, N, H, W, seed=SEED + 2, device="cpu"):
    """B x N x 5 x H x W surrogate stage-2 output."""
    g = _gen(seed, device)
    y = torch.randn((B, N, 5, H, W), generator=g, device=device)
    y[:, :, 0] *= 2.0
    y[:, :, 1:] *= 0.5
    return y


def timesteps(B, N, device="cpu"):
    """B x N times k/(N+1), k = 1..N (N=7 -> k/8, N=31 -> k/32)."""
    t = torch.arange(1, N + 1, dtype=torch.float32, device=device) / float(N + 1)
    return t.view(1, N).expand(B, N).contiguous()


def random_timesteps(B, N=1, seed=SEED + 3, device="cpu"):
    """per-sample random k/8, k in [1,7], as the training reader draws them (default_reader.py:167)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    k = torch.randint(1, 8, (B, N), generator=g)
    return (k.float() / 8.0).to(device)


def pad32(n):
    """Reference padding rule: ceil to a multiple of 32 (evaluate_interpolation_results.py:89-90)."""
    return (n + 31) // 32 * 32


def images_u8(F_, h, w, seed=SEED, device="cpu", smooth=True):
    """F x h x w x 3 uint8 images (RGB byte order): the low-passed noise of `frames` quantised to 8 bits, as a decoded
    video frame is (the reference reads uint8 images, scripts/visualize_interpolation.py:61-72), or white noise."""
    g = _gen(seed, device)
    if not smooth:
        return torch.randint(0, 256, (F_, h, w, 3), dtype=torch.uint8, generator=g, device=device)
    x = torch.rand((F_, 3, h, w), generator=g, device=device)
    k = torch.full((3, 1, 5, 5), 1.0 / 25.0, device=device)
    for _ in range(4):
        x = F.conv2d(F.pad(x, (2, 2, 2, 2), mode="replicate"), k, groups=3)
    lo = x.amin(dim=(2, 3), keepdim=True)
    hi = x.amax(dim=(2, 3), keepdim=True)
    x = (x - lo) / (hi - lo + 1e-12)
    return (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
