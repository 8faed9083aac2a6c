"""Element-wise steps between the U-Nets' cuDNN convolutions (ssm_upsample2x_nhwc, ssm_bias_leaky_nhwc,
ssm_avgpool2_nhwc of include/ssm_b200.h), for channels-last activations in inference.

The convolutions stay on PyTorch / cuDNN.  What these replace are the ATen element-wise ops around them, which
take 60 % of a whole 1080p inference step (profiles/r01s_pipeline_profile.txt): the channels-last bilinear
upsampling (the `upsampleN` lambdas of scripts/models/flow_computation.py:92-134), the bias add + LeakyReLU(0.1)
of layers.conv (scripts/models/layers.py:21-33) and AvgPool2d(2) (layers.py:60-63).  Same arithmetic, same
operation order as the ATen ops; forward only -- `usable(x)` is False whenever autograd is recording, and the
callers in unets.py then use the stock torch ops.
"""
import ctypes

import torch

from . import _abi


def usable(x):
    """CUDA, channels-last dense, bf16/fp32, C a multiple of 8, and no autograd graph to extend."""
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] % 8 == 0
            and not (torch.is_grad_enabled() and x.requires_grad)
            and x.is_contiguous(memory_format=torch.channels_last))


def _nhwc_empty(M, C, H, W, like):
    return torch.empty((M, C, H, W), dtype=like.dtype, device=like.device, memory_format=torch.channels_last)


def upsample2x_cat(parts):
    """F.interpolate(torch.cat(parts, dim=1), size=(2H, 2W), mode="bilinear", align_corners=False) without the
    concatenated intermediate: every part is upsampled straight into its channel slice of the result."""
    x0 = parts[0]
    M, _, H, W = x0.shape
    C = sum(p.shape[1] for p in parts)
    out = _nhwc_empty(M, C, 2 * H, 2 * W, x0)
    L = _abi.lib()
    esz = out.element_size()
    off = 0
    with torch.cuda.device_of(x0):
        for p in parts:
            if p.shape[0] != M or p.shape[2:] != x0.shape[2:] or p.dtype != x0.dtype:
                raise RuntimeError("upsample2x_cat: parts must share batch, size and dtype")
            _abi.check(L.ssm_upsample2x_nhwc(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(out.data_ptr() + off * esz),
                                             M, H, W, p.shape[1], C, _abi.dtype_code(p), _abi.stream_ptr(p.device)),
                       "ssm_upsample2x_nhwc")
            off += p.shape[1]
    return out


def bias_leaky_(y, bias_f32, slope=0.1):
    """y <- leaky_relu(y + bias, slope) in place; bias_f32: fp32 CUDA tensor of C values."""
    M, C, H, W = y.shape
    with torch.cuda.device_of(y):
        _abi.check(_abi.lib().ssm_bias_leaky_nhwc(ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(bias_f32.data_ptr()),
                                                  M * H * W, C, float(slope), _abi.dtype_code(y), _abi.stream_ptr(y.device)),
                   "ssm_bias_leaky_nhwc")
    return y


def avgpool2(x):
    """AvgPool2d(2) of a channels-last tensor with even H and W."""
    M, C, H, W = x.shape
    out = _nhwc_empty(M, C, H // 2, W // 2, x)
    with torch.cuda.device_of(x):
        _abi.check(_abi.lib().ssm_avgpool2_nhwc(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                M, H // 2, W // 2, C, _abi.dtype_code(x), _abi.stream_ptr(x.device)),
                   "ssm_avgpool2_nhwc")
    return out
