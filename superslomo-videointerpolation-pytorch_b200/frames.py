"""The steps either side of the synthesis path, over
ssm_frames_from_u8 / ssm_frames_to_u8:

  frames_from_u8   uint8 H x W x 3 images (cv2 BGR or RGB) -> normalised frames padded to a multiple of
                   32, planar NCHW and/or the RGBx copy the gather kernels read, in one launch
                   (reference: scripts/visualize_interpolation.py:61-88 load_batch + :257-262
                   normalize_tensor; scripts/utils/dataloaders/augmentations.py:181-190 Normalize +
                   default_reader.py:266-271 EvalPad)
  frames_to_u8     planar frames -> crop, de-normalise, uint8 H x W x 3
                   (reference: scripts/evaluate_interpolation_results.py:143-163, 192-202;
                   scripts/visualize_interpolation.py:221-232, 264-268)

The normalisation of a byte is a 3 x 256 table; `normalisation_lut` fills it with the reference's own
expression, evaluated by torch on the device whose bit pattern is wanted ("visualize": fp32 torch ops,
CPU or CUDA; "reader": the data loader's float64 numpy arithmetic rounded to fp32).
"""
import ctypes

import numpy as np
import torch

from . import _abi
from .synthetic import PIXEL_MEAN, PIXEL_STD


def normalisation_lut(mean=PIXEL_MEAN, std=PIXEL_STD, divisor=255.0, style="visualize", device="cuda"):
    """3 x 256 fp32 table on `device`: lut[c, v] = normalised value of byte v in channel c (R, G, B)."""
    if style == "reader":       # augmentations.py:187-189 on a uint8 numpy array: float64 arithmetic
        v = np.arange(256, dtype=np.uint8)[None, :]
        m, s = np.asarray(mean, dtype=np.float64)[:, None], np.asarray(std, dtype=np.float64)[:, None]
        lut = torch.from_numpy(((v / divisor - m) / s)).float()
        return lut.to(device).contiguous()
    if style != "visualize":
        raise ValueError("style must be 'visualize' or 'reader'")
    dev = torch.device(device)  # visualize_interpolation.py:257-260, same ops on the chosen device
    v = torch.arange(256, dtype=torch.float32, device=dev).view(1, 256)
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1).to(dev)
    s = torch.tensor(std, dtype=torch.float32).view(3, 1).to(dev)
    return ((v / divisor - m) / s).contiguous()


def center_padding(h_in, w_in, multiple=32):
    """(H, W, top, left) of the reference's padding rule: ceil to a multiple of 32, the smaller half
    of the padding first (visualize_interpolation.py:76-85; evaluate_interpolation_results.py:89-93)."""
    H = (h_in + multiple - 1) // multiple * multiple
    W = (w_in + multiple - 1) // multiple * multiple
    return H, W, (H - h_in) // 2, (W - w_in) // 2


def frames_from_u8(images, order="bgr", pad_mode="before", lut=None, dtype=torch.float32, want_planar=True,
                   want_rgbx=False, multiple=32, pad_values=None):
    """images: F x H_in x W_in x 3 uint8 CUDA tensor (rows and pixels dense).  Returns
    (planar F x 3 x H x W or None, rgbx F x H x W x 4 or None, (top, left)).

    pad_mode "before": pad with byte 0 and normalise everything (visualize_interpolation.py:87 then :137);
    "after": normalise, then zero-pad (default_reader.py:266-271).
    pad_values: the three fill values of the padding as host floats; when None and pad_mode is "before" they are
    read back from the table (lut[:, 0]), which synchronises -- pass them (e.g. `lut[:, 0].tolist()` computed
    once) to keep the call asynchronous and capturable into a CUDA graph."""
    if not images.is_cuda or images.dtype != torch.uint8:
        raise RuntimeError("frames_from_u8 needs a uint8 CUDA tensor (no CPU fallback); got %s on %s"
                           % (images.dtype, images.device))
    if images.dim() != 4 or images.shape[-1] != 3 or images.stride(-1) != 1 or images.stride(-2) != 3:
        raise RuntimeError("frames_from_u8: expected F x H x W x 3 with dense pixels, got %s / strides %s"
                           % (tuple(images.shape), images.stride()))
    F, h_in, w_in, _ = images.shape
    H, W, top, left = center_padding(h_in, w_in, multiple)
    dev = images.device
    if lut is None:
        lut = normalisation_lut(device=dev)
    lut = lut.to(device=dev, dtype=torch.float32).contiguous()
    if pad_mode not in ("before", "after"):
        raise ValueError("pad_mode must be 'before' or 'after'")
    if pad_values is not None:
        pad = [float(x) for x in pad_values]
        if len(pad) != 3:
            raise ValueError("pad_values must hold three floats (R, G, B)")
    elif pad_mode == "before":
        pad = lut[:, 0].cpu()
    else:
        pad = torch.zeros(3)
    pad_arr = (ctypes.c_float * 3)(*[float(x) for x in pad])
    planar = torch.empty((F, 3, H, W), dtype=dtype, device=dev) if want_planar else None
    rgbx = torch.empty((F, H, W, 4), dtype=dtype, device=dev) if want_rgbx else None
    with torch.cuda.device(dev):
        rc = _abi.lib().ssm_frames_from_u8(
            ctypes.c_void_p(images.data_ptr()), images.stride(0), images.stride(1), 1 if order.lower() == "bgr" else 0,
            F, h_in, w_in, H, W, top, left, ctypes.c_void_p(lut.data_ptr()), pad_arr,
            _abi.ref(_abi.desc(planar, False)), ctypes.c_void_p(rgbx.data_ptr()) if rgbx is not None else None,
            _abi.dtype_code(planar if planar is not None else rgbx), _abi.stream_ptr(dev))
    _abi.check(rc, "ssm_frames_from_u8")
    return planar, rgbx, (top, left)


def frames_to_u8(frames, top=0, left=0, h_out=None, w_out=None, mean=PIXEL_MEAN, std=PIXEL_STD, scale=255.0,
                 order="rgb", saturate=False):
    """frames: F x 3 x H x W CUDA tensor (fp32 or bf16) -> F x h_out x w_out x 3 uint8 CUDA tensor:
    crop, (x * std + mean) * scale, numpy-style astype(uint8) (wraps out-of-range values like the
    reference; saturate=True clamps to [0, 255] instead)."""
    if not frames.is_cuda:
        raise RuntimeError("frames_to_u8 runs on CUDA tensors only (no CPU fallback)")
    frames = _abi.dense_planes(frames)
    F, C, H, W = frames.shape
    if C != 3:
        raise RuntimeError("frames_to_u8: expected F x 3 x H x W, got %s" % (tuple(frames.shape),))
    h_out = H - top if h_out is None else h_out
    w_out = W - left if w_out is None else w_out
    out = torch.empty((F, h_out, w_out, 3), dtype=torch.uint8, device=frames.device)
    m = (ctypes.c_float * 3)(*[float(x) for x in mean])
    s = (ctypes.c_float * 3)(*[float(x) for x in std])
    with torch.cuda.device(frames.device):
        rc = _abi.lib().ssm_frames_to_u8(
            _abi.ref(_abi.desc(frames, False)), F, H, W, top, left, h_out, w_out, m, s, float(scale),
            1 if order.lower() == "bgr" else 0, 1 if saturate else 0, ctypes.c_void_p(out.data_ptr()),
            out.stride(0), out.stride(1), _abi.dtype_code(frames), _abi.stream_ptr(frames.device))
    _abi.check(rc, "ssm_frames_to_u8")
    return out
