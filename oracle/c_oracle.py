"""ctypes front-end of the C oracle (oracle/ssm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

All functions take and return contiguous fp32 CPU torch tensors (NCHW) and follow the
reference call surface:
  warp(x, flo)                                     scripts/models/layers.py:73
  compute_inputs(img6, flow4, t)                   scripts/models/flow_interpolation.py:338
  compute_output_image(img6, in16, out5, t)        scripts/models/flow_interpolation.py:394
plus explicit backward functions (the reference relies on autograd).
`t` is a tensor with one value per sample (any shape with B elements, e.g. B x 1 x 1 x 1).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssm_oracle.so")

COORD_DIV = 0  # true division   -- torch CPU
COORD_RCP = 1  # reciprocal-mul  -- torch CUDA (Python-scalar divisor)


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "ssm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.c_void_p
        i = ctypes.c_int
        _lib.ssm_oracle_warp_fwd.argtypes = [fp, fp, fp, i, i, i, i, i]
        _lib.ssm_oracle_warp_bwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i, i]
        _lib.ssm_oracle_flow_pack_fwd.argtypes = [fp, fp, fp, fp, i, i, i, i]
        _lib.ssm_oracle_flow_pack_bwd.argtypes = [fp, fp, fp, fp, fp, fp, i, i, i, i]
        _lib.ssm_oracle_fuse_fwd.argtypes = [fp, fp, fp, fp, fp, i, i, i, i]
        _lib.ssm_oracle_fuse_bwd.argtypes = [fp, fp, fp, fp, fp, fp, fp, fp, i, i, i, i]
        _lib.ssm_oracle_num_threads.restype = i
        for name in ("warp_fwd", "warp_bwd", "flow_pack_fwd", "flow_pack_bwd", "fuse_fwd", "fuse_bwd"):
            getattr(_lib, "ssm_oracle_" + name).restype = None
    return _lib


def num_threads():
    return int(lib().ssm_oracle_num_threads())


def _c(x):
    assert x.device.type == "cpu", "the oracle runs on CPU tensors"
    return x.detach().to(torch.float32).contiguous()


def _p(x):
    return ctypes.c_void_p(x.data_ptr()) if x is not None else ctypes.c_void_p(0)


def _tvec(t, B):
    t = _c(torch.as_tensor(t)).reshape(-1)
    if t.numel() == 1 and B > 1:
        t = t.expand(B).contiguous()
    assert t.numel() == B, "t needs one value per sample"
    return t


def warp(x, flo, coord_mode=COORD_DIV):
    x, flo = _c(x), _c(flo)
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    lib().ssm_oracle_warp_fwd(_p(x), _p(flo), _p(out), B, C, H, W, coord_mode)
    return out


def warp_backward(grad_out, x, flo, coord_mode=COORD_DIV, need_img=True, need_flow=True):
    grad_out, x, flo = _c(grad_out), _c(x), _c(flo)
    B, C, H, W = x.shape
    gx = torch.empty_like(x) if need_img else None
    gf = torch.empty_like(flo) if need_flow else None
    lib().ssm_oracle_warp_bwd(_p(grad_out), _p(x), _p(flo), _p(gx), _p(gf), B, C, H, W, coord_mode)
    return gx, gf


def compute_inputs(img6, flow4, t, coord_mode=COORD_DIV):
    img6, flow4 = _c(img6), _c(flow4)
    B, _, H, W = img6.shape
    t = _tvec(t, B)
    out = torch.empty(B, 16, H, W)
    lib().ssm_oracle_flow_pack_fwd(_p(img6), _p(flow4), _p(t), _p(out), B, H, W, coord_mode)
    return out


def compute_inputs_backward(g16, img6, flow4, t, coord_mode=COORD_DIV, need_img=True):
    g16, img6, flow4 = _c(g16), _c(img6), _c(flow4)
    B, _, H, W = img6.shape
    t = _tvec(t, B)
    gflow = torch.empty_like(flow4)
    gimg = torch.empty_like(img6) if need_img else None
    lib().ssm_oracle_flow_pack_bwd(_p(g16), _p(img6), _p(flow4), _p(t), _p(gflow), _p(gimg), B, H, W, coord_mode)
    return gimg, gflow


def compute_output_image(img6, in16, out5, t, coord_mode=COORD_DIV):
    img6, in16, out5 = _c(img6), _c(in16), _c(out5)
    B, _, H, W = img6.shape
    t = _tvec(t, B)
    out = torch.empty(B, 3, H, W)
    lib().ssm_oracle_fuse_fwd(_p(img6), _p(in16), _p(out5), _p(t), _p(out), B, H, W, coord_mode)
    return out


def compute_output_image_backward(g3, img6, in16, out5, t, coord_mode=COORD_DIV, need_img=True):
    g3, img6, in16, out5 = _c(g3), _c(img6), _c(in16), _c(out5)
    B, _, H, W = img6.shape
    t = _tvec(t, B)
    gout5 = torch.empty_like(out5)
    gin16 = torch.empty_like(in16)
    gimg = torch.empty_like(img6) if need_img else None
    lib().ssm_oracle_fuse_bwd(_p(g3), _p(img6), _p(in16), _p(out5), _p(t), _p(gout5), _p(gin16), _p(gimg),
                              B, H, W, coord_mode)
    return gimg, gin16, gout5
