"""Drop-in for the hot-path part of the reference's scripts/models/layers.py.

`warp(x, flo)` keeps the reference signature (layers.py:73) and runs the sm_100a kernel through
the C ABI.  `conv` and `avg_pool` are kept call-compatible because the reference's U-Nets import
them from the same module (flow_interpolation.py:9); they are plain torch/cuDNN layers and are
not part of the rebuilt path.
"""
import torch.nn as nn

from . import functional as F_ssm


def warp(x, flo, packed=None):
    """Backward warp of x (B x C x H x W) by flow flo (B x 2 x H x W), bilinear, zeros outside,
    corners aligned -- same results as reference layers.warp (layers.py:73-120), differentiable
    in x and flo.  CUDA tensors only.  packed: optional RGBx copy of a 3-channel x (ssm_b200.pack_image) shared by
    several warps of the same image."""
    return F_ssm.warp(x, flo, packed=packed)


def conv(in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1):
    """Conv2d + LeakyReLU(0.1), as the reference's layers.conv (layers.py:21-33)."""
    layer = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                      dilation=dilation, bias=True)
    return nn.Sequential(layer, nn.LeakyReLU(0.1, inplace=True))


def avg_pool(kernel_size=2, stride=None, padding=0):
    """AvgPool2d as the reference's layers.avg_pool (layers.py:60-63)."""
    return nn.AvgPool2d(kernel_size, stride, padding, ceil_mode=False, count_include_pad=True)
