"""Differentiable torch entry points over the C ABI (include/ssm_b200.h).

The reference has no hand-written backward: its path is differentiable because it is made of
torch ops (scripts/models/layers.py:73-120, scripts/models/flow_interpolation.py:338-429).  Here
each fused kernel pair is a torch.autograd.Function with gradients for every floating input
except t, which is what autograd derives for the reference.

All functions require CUDA tensors; nothing here computes on the CPU.
"""
import ctypes
import os

import torch

from . import _abi

_COORD_MODES = {"cpu": _abi.COORD_DIV, "div": _abi.COORD_DIV, "cuda": _abi.COORD_RCP, "rcp": _abi.COORD_RCP}
_default_coord_mode = _abi.COORD_DIV


def set_coord_mode(mode):
    """Select which reference bit-pattern the sampling coordinates follow (SURVEY.md finding 3b).

    "cpu" / "div": IEEE division by max(W-1,1)  -- bit-matches the reference run on CPU (default)
    "cuda" / "rcp": multiply by fp32 1/(W-1)     -- bit-matches the reference run on CUDA (ATen)
    """
    global _default_coord_mode
    _default_coord_mode = _resolve_mode(mode)


def get_coord_mode():
    return _default_coord_mode


def _resolve_mode(mode):
    if mode is None:
        return _default_coord_mode
    if isinstance(mode, str):
        return _COORD_MODES[mode.lower()]
    if mode in (_abi.COORD_DIV, _abi.COORD_RCP):
        return int(mode)
    raise ValueError("unknown coord mode %r" % (mode,))


def _same(*ts):
    t0 = ts[0]
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("ssm_b200 runs on CUDA tensors only -- there is no CPU fallback "
                               "(got a tensor on %s)" % t.device)
    for t in ts[1:]:
        if t.dtype != t0.dtype or t.device != t0.device:
            raise RuntimeError("ssm_b200: all tensors of a call must share dtype and device "
                               "(got %s/%s and %s/%s)" % (t0.dtype, t0.device, t.dtype, t.device))


def _workspace(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def _out_buffer(out, shape, like, what):
    """Caller-owned result buffer (`out=`): validated, or allocated when None.  A serving loop that
    owns its buffers makes no allocator calls in steady state (the C ABI itself never allocates)."""
    if out is None:
        return torch.empty(shape, dtype=like.dtype, device=like.device)
    if tuple(out.shape) != tuple(shape) or out.dtype != like.dtype or out.device != like.device \
            or not out.is_contiguous():
        raise RuntimeError("%s: out= must be a contiguous %s %s tensor on %s (got %s %s on %s)"
                           % (what, tuple(shape), like.dtype, like.device, tuple(out.shape), out.dtype, out.device))
    return out


def _no_grad_for_out(out, what, *inputs):
    if out is not None and torch.is_grad_enabled() and any(x.requires_grad for x in inputs):
        raise RuntimeError("%s: out= is for inference; it cannot be combined with inputs that require grad" % what)


# How a DEVICE-side t is range-checked (the reference asserts 0 < t < 1, scripts/utils/validators.py:9-11; a host-side
# t is always checked).  Reading a device tensor would synchronise every call, so the default is "off"; "flag" counts
# violations in a device counter (one tiny kernel per call, read with t_violations() whenever the caller synchronises
# anyway); "assert" uses torch._assert_async (a violation traps the kernel: for debugging).  SSM_B200_CHECK_T=flag|assert.
_T_CHECK = os.environ.get("SSM_B200_CHECK_T", "off")
_t_violations = {}


def set_device_t_check(mode):
    """mode: "off" | "flag" | "assert".  Returns the previous mode."""
    global _T_CHECK
    if mode not in ("off", "flag", "assert"):
        raise ValueError("set_device_t_check: mode must be off, flag or assert")
    previous, _T_CHECK = _T_CHECK, mode
    return previous


def t_violations(device=None, reset=True):
    """number of device-side t values outside (0, 1) seen since the last reset (synchronises)"""
    total = 0
    for dev, counter in list(_t_violations.items()):
        if device is None or torch.device(device) == dev:
            total += int(counter.item())
            if reset:
                counter.zero_()
    return total


def _t_vector(t, count, device):
    """t as `count` contiguous fp32 values on `device` (t is never differentiated)."""
    t = torch.as_tensor(t).detach()
    if not t.is_cuda and t.numel() > 0:
        # the reference asserts 0 < t < 1 (scripts/utils/validators.py:9-11); a host-side t is checked here
        lo, hi = float(t.min()), float(t.max())
        if not (0.0 < lo and hi < 1.0):
            raise AssertionError("t must satisfy 0 < t < 1 (got values in [%g, %g])" % (lo, hi))
    elif t.is_cuda and t.numel() > 0 and _T_CHECK != "off" and not torch.cuda.is_current_stream_capturing():
        inside = (t > 0) & (t < 1)
        if _T_CHECK == "assert":
            torch._assert_async(inside.all(), "t must satisfy 0 < t < 1")
        else:
            counter = _t_violations.get(t.device)
            if counter is None:
                counter = _t_violations[t.device] = torch.zeros((), dtype=torch.int64, device=t.device)
            counter += (~inside).sum()
    t = t.to(device=device, dtype=torch.float32).reshape(-1)
    if t.numel() == 1 and count > 1:
        t = t.expand(count)
    if t.numel() != count:
        raise RuntimeError("ssm_b200: t has %d values, expected %d (one per pair and timestep)" % (t.numel(), count))
    return t.contiguous()


def _empty_result(shape, out, *inputs):
    """An empty batch (B = 0 or N = 0): the reference's torch ops return an empty tensor whose gradients are empty
    too (scripts/models/layers.py:73-120 on a 0 x C x H x W input); nothing to launch.  The result stays attached
    to the inputs' graph."""
    if out is not None:
        return out
    like = inputs[0]
    res = torch.zeros(shape, dtype=like.dtype, device=like.device)
    for t in inputs:
        if isinstance(t, torch.Tensor) and t.requires_grad:
            res = res + t.sum() * 0
    return res


class _NoCtx:
    """stand-in for the autograd context when a forward is called directly with out= (inference)"""

    def save_for_backward(self, *tensors):
        pass


# ---------------------------------------------------------------------------------------------
class _Warp(torch.autograd.Function):
    """layers.warp(x, flo) -- reference scripts/models/layers.py:73-120"""

    @staticmethod
    def forward(ctx, x, flo, mode, packed):
        _same(x, flo)
        x, flo = _abi.dense_planes(x), _abi.dense_planes(flo)
        B, C, H, W = x.shape
        if flo.shape != (B, 2, H, W):
            raise RuntimeError("warp: flo must be B x 2 x H x W, got %s for x %s" % (tuple(flo.shape), tuple(x.shape)))
        if packed is not None and (C != 3 or packed.shape != (B, H, W, 4) or packed.dtype != x.dtype
                                   or packed.device != x.device or not packed.is_contiguous()):
            raise RuntimeError("warp: packed= must be the contiguous %s %s RGBx copy of a 3-channel x (pack_image)"
                               % ((B, H, W, 4), x.dtype))
        out = torch.empty((B, C, H, W), dtype=x.dtype, device=x.device)
        L = _abi.lib()
        with torch.cuda.device(x.device):
            if packed is not None:
                rc = L.ssm_warp_fwd_packed(ctypes.c_void_p(packed.data_ptr()), _abi.ref(_abi.desc(flo, False)),
                                           _abi.ref(_abi.desc(out, False)), B, H, W, _abi.dtype_code(x), mode,
                                           _abi.stream_ptr(x.device))
            else:
                rc = L.ssm_warp_fwd(_abi.ref(_abi.desc(x, False)), _abi.ref(_abi.desc(flo, False)),
                                    _abi.ref(_abi.desc(out, False)), B, C, H, W, _abi.dtype_code(x), mode,
                                    _abi.stream_ptr(x.device))
        _abi.check(rc, "ssm_warp_fwd")
        ctx.save_for_backward(x, flo, packed)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, gout):
        x, flo, packed = ctx.saved_tensors
        B, C, H, W = x.shape
        need_x, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_f):
            return None, None, None, None
        gout = _abi.dense_planes(gout.to(x.dtype))
        gx = torch.empty_like(x, memory_format=torch.contiguous_format) if need_x else None
        gf = torch.empty_like(flo, memory_format=torch.contiguous_format) if need_f else None
        L = _abi.lib()
        ws, ws_ptr, ws_bytes = None, None, 0
        if need_x:
            ws_bytes = L.ssm_warp_bwd_workspace_bytes(B, C, H, W)
            ws = _workspace(ws_bytes, x.device)
            ws_ptr = ctypes.c_void_p(ws.data_ptr())
        with torch.cuda.device(x.device):
            if packed is not None:
                rc = L.ssm_warp_bwd_packed(_abi.ref(_abi.desc(gout, False)), ctypes.c_void_p(packed.data_ptr()),
                                           _abi.ref(_abi.desc(flo, False)), _abi.ref(_abi.desc(gx, False)),
                                           _abi.ref(_abi.desc(gf, False)), B, H, W, _abi.dtype_code(x), ctx.mode,
                                           ws_ptr, ws_bytes, _abi.stream_ptr(x.device))
            else:
                rc = L.ssm_warp_bwd(_abi.ref(_abi.desc(gout, False)), _abi.ref(_abi.desc(x, False)),
                                    _abi.ref(_abi.desc(flo, False)), _abi.ref(_abi.desc(gx, False)),
                                    _abi.ref(_abi.desc(gf, False)), B, C, H, W, _abi.dtype_code(x), ctx.mode,
                                    ws_ptr, ws_bytes, _abi.stream_ptr(x.device))
        _abi.check(rc, "ssm_warp_bwd")
        return gx, gf, None, None


def warp(x, flo, coord_mode=None, packed=None):
    """Backward-warp image x (B x C x H x W) by flow flo (B x 2 x H x W; channel 0 horizontal).  packed: the RGBx
    copy of a 3-channel x (pack_image), for callers that warp the same image by several flows: one 16-byte gather
    per bilinear tap instead of three 4-byte ones; gradients still flow to x."""
    if x.shape[0] == 0 and x.is_cuda:
        return _empty_result(x.shape, None, x, flo)
    return _Warp.apply(x, flo, _resolve_mode(coord_mode), packed)


def pack_image(x, out=None):
    """RGBx re-layout of B x 3 x H x W images (ssm_pack_image) -> B x H x W x 4, for warp(..., packed=).  Not
    differentiable (a staging copy; the gradient of warp reaches x itself)."""
    _same(x)
    x = _abi.dense_planes(x.detach())
    B, C, H, W = x.shape
    if C != 3:
        raise RuntimeError("pack_image: expected B x 3 x H x W, got %s" % (tuple(x.shape),))
    packed = _out_buffer(out, (B, H, W, 4), x, "pack_image")
    with torch.cuda.device(x.device):
        rc = _abi.lib().ssm_pack_image(_abi.ref(_abi.desc(x, False)), ctypes.c_void_p(packed.data_ptr()),
                                       B, H, W, _abi.dtype_code(x), _abi.stream_ptr(x.device))
    _abi.check(rc, "ssm_pack_image")
    return packed


# ---------------------------------------------------------------------------------------------
def pack_frames(img6, out=None):
    """RGBx re-layout of a batch of frame pairs (ssm_pack_frames): B x 6 x H x W planar ->
    B x 2 x H x W x 4 pixel-interleaved, so each bilinear tap of the gathers is one memory request.
    Optional: flow_pack / fuse build it themselves when they are given N >= 2 timesteps; build it
    once here and pass it to both to share it.  Not differentiable (it is a staging copy)."""
    _same(img6)
    img6 = _abi.dense_planes(img6.detach())
    B, C6, H, W = img6.shape
    if C6 != 6:
        raise RuntimeError("pack_frames: expected B x 6 x H x W, got %s" % (tuple(img6.shape),))
    packed = _out_buffer(out, (B, 2, H, W, 4), img6, "pack_frames")
    with torch.cuda.device(img6.device):
        rc = _abi.lib().ssm_pack_frames(_abi.ref(_abi.desc(img6, False)), ctypes.c_void_p(packed.data_ptr()),
                                        B, H, W, _abi.dtype_code(img6), _abi.stream_ptr(img6.device))
    _abi.check(rc, "ssm_pack_frames")
    return packed


def _packed_ptr(packed, img6):
    if packed is None:
        return None
    B, _, H, W = img6.shape
    if packed.shape != (B, 2, H, W, 4) or packed.dtype != img6.dtype or packed.device != img6.device \
            or not packed.is_contiguous():
        raise RuntimeError("packed frames do not match img_tensor (expected contiguous %s %s)"
                           % ((B, 2, H, W, 4), img6.dtype))
    return ctypes.c_void_p(packed.data_ptr())


_AUTO_PACK_MIN_TIMESTEPS = 2


class _FlowPack(torch.autograd.Function):
    """compute_inputs for N timesteps -- reference scripts/models/flow_interpolation.py:338-372"""

    @staticmethod
    def forward(ctx, img6, flow4, tvec, N, mode, packed, out=None):
        _same(img6, flow4)
        img6, flow4 = _abi.dense_planes(img6), _abi.dense_planes(flow4)
        B, C6, H, W = img6.shape
        if C6 != 6 or flow4.shape != (B, 4, H, W):
            raise RuntimeError("compute_inputs: expected img B x 6 x H x W and flow B x 4 x H x W, got %s and %s"
                               % (tuple(img6.shape), tuple(flow4.shape)))
        if packed is None and N >= _AUTO_PACK_MIN_TIMESTEPS:
            packed = pack_frames(img6)
        out = _out_buffer(out, (B, N, 16, H, W), img6, "flow_pack")
        with torch.cuda.device(img6.device):
            rc = _abi.lib().ssm_flow_pack_fwd(_abi.ref(_abi.desc(img6, False)), _packed_ptr(packed, img6),
                                              _abi.ref(_abi.desc(flow4, False)),
                                              ctypes.c_void_p(tvec.data_ptr()), _abi.ref(_abi.desc(out, True)),
                                              B, N, H, W, _abi.dtype_code(img6), mode, _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_flow_pack_fwd")
        ctx.save_for_backward(img6, flow4, tvec, packed)
        ctx.N, ctx.mode = N, mode
        return out

    @staticmethod
    def backward(ctx, g16):
        img6, flow4, tvec, packed = ctx.saved_tensors
        B, _, H, W = img6.shape
        N = ctx.N
        need_i, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_i or need_f):
            return None, None, None, None, None, None
        g16 = _abi.dense_planes(g16.to(img6.dtype))
        gi = torch.empty_like(img6, memory_format=torch.contiguous_format) if need_i else None
        gf = torch.empty_like(flow4, memory_format=torch.contiguous_format) if need_f else None
        L = _abi.lib()
        ws, ws_ptr, ws_bytes = None, None, 0
        if need_i:
            ws_bytes = L.ssm_flow_pack_bwd_workspace_bytes(B, N, H, W)
            ws = _workspace(ws_bytes, img6.device)
            ws_ptr = ctypes.c_void_p(ws.data_ptr())
        with torch.cuda.device(img6.device):
            rc = L.ssm_flow_pack_bwd(_abi.ref(_abi.desc(g16, True)), _abi.ref(_abi.desc(img6, False)),
                                     _packed_ptr(packed, img6),
                                     _abi.ref(_abi.desc(flow4, False)), ctypes.c_void_p(tvec.data_ptr()),
                                     _abi.ref(_abi.desc(gf, False)), _abi.ref(_abi.desc(gi, False)),
                                     B, N, H, W, _abi.dtype_code(img6), ctx.mode, ws_ptr, ws_bytes,
                                     _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_flow_pack_bwd")
        return gi, gf, None, None, None, None


def flow_pack(img6, flow4, t, n_timesteps=1, coord_mode=None, packed=None, out=None):
    """Stage-2 input for n_timesteps intermediate times of every pair: B x N x 16 x H x W.

    t holds B*N values (t[b, n]); a single value is broadcast.  packed: optional result of
    pack_frames(img6) to share the RGBx copy between flow_pack and fuse.  out: optional caller-owned
    result buffer (inference only)."""
    B = img6.shape[0]
    if B * int(n_timesteps) == 0 and img6.is_cuda:
        return _empty_result((B, int(n_timesteps), 16) + tuple(img6.shape[2:]), out, img6, flow4)
    tvec = _t_vector(t, B * n_timesteps, img6.device)
    if out is not None:
        _no_grad_for_out(out, "flow_pack", img6, flow4)
        with torch.no_grad():
            return _FlowPack.forward(_NoCtx(), img6, flow4, tvec, int(n_timesteps), _resolve_mode(coord_mode), packed, out)
    return _FlowPack.apply(img6, flow4, tvec, int(n_timesteps), _resolve_mode(coord_mode), packed)


def flow_pack_channels_last(img6, flow4, t, n_timesteps=1, dtype=None, coord_mode=None, packed=None, out=None):
    """compute_inputs for N timesteps, written in the layout (and dtype) the stage-2 U-Net consumes when it
    runs channels-last, optionally under bf16 autocast (ssm_flow_pack_fwd_nhwc).

    Returns a B x N x 16 x H x W tensor whose memory is B x N x H x W x 16 (`x.view(B*N, 16, H, W)` is a
    torch.channels_last tensor: conv1a takes it without a layout or dtype conversion pass).  dtype: None
    (= the frames' dtype) or torch.bfloat16; the values are those of flow_pack rounded once to dtype.
    Inference only (the training loop uses flow_pack, whose backward exists).  out: optional caller-owned
    result buffer with exactly that shape and strides."""
    _same(img6, flow4)
    if torch.is_grad_enabled() and (img6.requires_grad or flow4.requires_grad):
        raise RuntimeError("flow_pack_channels_last is inference-only; use flow_pack for a differentiable result")
    img6, flow4 = _abi.dense_planes(img6.detach()), _abi.dense_planes(flow4.detach())
    B, C6, H, W = img6.shape
    N = int(n_timesteps)
    if C6 != 6 or flow4.shape != (B, 4, H, W):
        raise RuntimeError("compute_inputs: expected img B x 6 x H x W and flow B x 4 x H x W, got %s and %s"
                           % (tuple(img6.shape), tuple(flow4.shape)))
    dtype = img6.dtype if dtype is None else dtype
    if dtype not in (torch.float32, torch.bfloat16) or (img6.dtype == torch.bfloat16 and dtype != torch.bfloat16):
        raise TypeError("flow_pack_channels_last: %s frames cannot be written as %s" % (img6.dtype, dtype))
    tvec = _t_vector(t, B * N, img6.device)
    if packed is None and N >= _AUTO_PACK_MIN_TIMESTEPS:
        packed = pack_frames(img6)
    if out is None:
        out = torch.empty((B, N, H, W, 16), dtype=dtype, device=img6.device).permute(0, 1, 4, 2, 3)
    elif tuple(out.shape) != (B, N, 16, H, W) or out.dtype != dtype or out.device != img6.device \
            or out.stride() != (N * 16 * H * W, 16 * H * W, 1, 16 * W, 16):
        raise RuntimeError("flow_pack_channels_last: out= must be a %s %s tensor stored as B x N x H x W x 16"
                           % ((B, N, 16, H, W), dtype))
    with torch.cuda.device(img6.device):
        rc = _abi.lib().ssm_flow_pack_fwd_nhwc(
            _abi.ref(_abi.desc(img6, False)), _packed_ptr(packed, img6), _abi.ref(_abi.desc(flow4, False)),
            ctypes.c_void_p(tvec.data_ptr()), ctypes.c_void_p(out.data_ptr()), B, N, H, W, _abi.dtype_code(img6),
            _abi.DTYPE_BF16 if dtype == torch.bfloat16 else _abi.DTYPE_F32, _resolve_mode(coord_mode),
            _abi.stream_ptr(img6.device))
    _abi.check(rc, "ssm_flow_pack_fwd_nhwc")
    return out


# ---------------------------------------------------------------------------------------------
class _Fuse(torch.autograd.Function):
    """extract_outputs + compute_output_image for N timesteps -- flow_interpolation.py:374-429"""

    @staticmethod
    def forward(ctx, img6, in16, out5, tvec, mode, packed, out=None):
        _same(img6, in16, out5)
        img6, in16, out5 = _abi.dense_planes(img6), _abi.dense_planes(in16), _abi.dense_planes(out5)
        B, C6, H, W = img6.shape
        N = in16.shape[1]
        if C6 != 6 or in16.shape != (B, N, 16, H, W) or out5.shape != (B, N, 5, H, W):
            raise RuntimeError("compute_output_image: expected img B x 6, input B x N x 16, output B x N x 5 "
                               "(x H x W), got %s, %s, %s" % (tuple(img6.shape), tuple(in16.shape), tuple(out5.shape)))
        if packed is None and N >= _AUTO_PACK_MIN_TIMESTEPS:
            packed = pack_frames(img6)
        out = _out_buffer(out, (B, N, 3, H, W), img6, "fuse")
        flows4 = in16[:, :, 6:10]
        with torch.cuda.device(img6.device):
            rc = _abi.lib().ssm_fuse_fwd(_abi.ref(_abi.desc(img6, False)), _packed_ptr(packed, img6),
                                         _abi.ref(_abi.desc(flows4, True)),
                                         _abi.ref(_abi.desc(out5, True)), ctypes.c_void_p(tvec.data_ptr()),
                                         _abi.ref(_abi.desc(out, True)), B, N, H, W, _abi.dtype_code(img6), mode,
                                         _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_fuse_fwd")
        ctx.save_for_backward(img6, in16, out5, tvec, packed)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g3):
        img6, in16, out5, tvec, packed = ctx.saved_tensors
        B, _, H, W = img6.shape
        N = in16.shape[1]
        need_i, need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_i or need_x or need_y):
            return None, None, None, None, None, None
        g3 = _abi.dense_planes(g3.to(img6.dtype))
        gi = torch.empty_like(img6, memory_format=torch.contiguous_format) if need_i else None
        gy = torch.empty_like(out5, memory_format=torch.contiguous_format) if need_y else None
        gx = None
        gx_flows = None
        if need_x:
            # the gradient of the 16-channel tensor is zero outside channels 6:10
            gx = torch.zeros_like(in16, memory_format=torch.contiguous_format)
            gx_flows = gx[:, :, 6:10]
        L = _abi.lib()
        ws, ws_ptr, ws_bytes = None, None, 0
        if need_i:
            ws_bytes = L.ssm_fuse_bwd_workspace_bytes(B, N, H, W)
            ws = _workspace(ws_bytes, img6.device)
            ws_ptr = ctypes.c_void_p(ws.data_ptr())
        with torch.cuda.device(img6.device):
            rc = L.ssm_fuse_bwd(_abi.ref(_abi.desc(g3, True)), _abi.ref(_abi.desc(img6, False)),
                                _packed_ptr(packed, img6),
                                _abi.ref(_abi.desc(in16[:, :, 6:10], True)), _abi.ref(_abi.desc(out5, True)),
                                ctypes.c_void_p(tvec.data_ptr()), _abi.ref(_abi.desc(gy, True)),
                                _abi.ref(_abi.desc(gx_flows, True)), _abi.ref(_abi.desc(gi, False)),
                                B, N, H, W, _abi.dtype_code(img6), ctx.mode, ws_ptr, ws_bytes,
                                _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_fuse_bwd")
        return gi, gx, gy, None, None, None


def fuse(img6, in16, out5, t, coord_mode=None, packed=None, out=None):
    """Fused frames for every (pair, timestep): img6 B x 6, in16 B x N x 16, out5 B x N x 5 -> B x N x 3.
    out: optional caller-owned result buffer (inference only)."""
    B, N = in16.shape[0], in16.shape[1]
    if B * N == 0 and img6.is_cuda:
        return _empty_result((B, N, 3) + tuple(img6.shape[2:]), out, img6, in16, out5)
    tvec = _t_vector(t, B * N, img6.device)
    if out is not None:
        _no_grad_for_out(out, "fuse", img6, in16, out5)
        with torch.no_grad():
            return _Fuse.forward(_NoCtx(), img6, in16, out5, tvec, _resolve_mode(coord_mode), packed, out)
    return _Fuse.apply(img6, in16, out5, tvec, _resolve_mode(coord_mode), packed)


# ---------------------------------------------------------------------------------------------
class _FuseFlow(torch.autograd.Function):
    """compute_output_image for N timesteps with the estimated flows recomputed from flow_pred_tensor
    (ssm_fuse_flow_fwd/bwd): same results as _Fuse on input_tensor = compute_inputs(img, flow, t)."""

    @staticmethod
    def forward(ctx, img6, flow4, out5, tvec, mode, packed, out=None):
        _same(img6, flow4, out5)
        img6, flow4, out5 = _abi.dense_planes(img6), _abi.dense_planes(flow4), _abi.dense_planes(out5)
        B, C6, H, W = img6.shape
        if out5.dim() != 5 or C6 != 6 or flow4.shape != (B, 4, H, W) or out5.shape[0] != B \
                or out5.shape[2:] != (5, H, W):
            raise RuntimeError("fuse_from_flow: expected img B x 6, flow B x 4, output B x N x 5 (x H x W), "
                               "got %s, %s, %s" % (tuple(img6.shape), tuple(flow4.shape), tuple(out5.shape)))
        N = out5.shape[1]
        if packed is None and N >= _AUTO_PACK_MIN_TIMESTEPS:
            packed = pack_frames(img6)
        out = _out_buffer(out, (B, N, 3, H, W), img6, "fuse_from_flow")
        with torch.cuda.device(img6.device):
            rc = _abi.lib().ssm_fuse_flow_fwd(_abi.ref(_abi.desc(img6, False)), _packed_ptr(packed, img6),
                                              _abi.ref(_abi.desc(flow4, False)), _abi.ref(_abi.desc(out5, True)),
                                              ctypes.c_void_p(tvec.data_ptr()), _abi.ref(_abi.desc(out, True)),
                                              B, N, H, W, _abi.dtype_code(img6), mode, _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_fuse_flow_fwd")
        ctx.save_for_backward(img6, flow4, out5, tvec, packed)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g3):
        img6, flow4, out5, tvec, packed = ctx.saved_tensors
        B, _, H, W = img6.shape
        N = out5.shape[1]
        need_i, need_f, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_i or need_f or need_y):
            return None, None, None, None, None, None
        g3 = _abi.dense_planes(g3.to(img6.dtype))
        gi = torch.empty_like(img6, memory_format=torch.contiguous_format) if need_i else None
        gf = torch.empty_like(flow4, memory_format=torch.contiguous_format) if need_f else None
        gy = torch.empty_like(out5, memory_format=torch.contiguous_format) if need_y else None
        L = _abi.lib()
        ws, ws_ptr, ws_bytes = None, None, 0
        if need_i:
            ws_bytes = L.ssm_fuse_bwd_workspace_bytes(B, N, H, W)
            ws = _workspace(ws_bytes, img6.device)
            ws_ptr = ctypes.c_void_p(ws.data_ptr())
        with torch.cuda.device(img6.device):
            rc = L.ssm_fuse_flow_bwd(_abi.ref(_abi.desc(g3, True)), _abi.ref(_abi.desc(img6, False)),
                                     _packed_ptr(packed, img6), _abi.ref(_abi.desc(flow4, False)),
                                     _abi.ref(_abi.desc(out5, True)), ctypes.c_void_p(tvec.data_ptr()),
                                     _abi.ref(_abi.desc(gy, True)), _abi.ref(_abi.desc(gf, False)),
                                     _abi.ref(_abi.desc(gi, False)), B, N, H, W, _abi.dtype_code(img6), ctx.mode,
                                     ws_ptr, ws_bytes, _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_fuse_flow_bwd")
        return gi, gf, gy, None, None, None


def _fuse_from_flow_mixed(img6, flow4, out5, tvec, mode, packed, out):
    """fp32 frames and flows with the U-Net output as bf16 autocast leaves it (ssm_fuse_flow_fwd_mixed):
    equals fuse_from_flow(img6, flow4, out5.float(), t) without materialising out5.float().  Inference only."""
    _same(img6, flow4)
    if torch.is_grad_enabled() and (img6.requires_grad or flow4.requires_grad or out5.requires_grad):
        raise RuntimeError("fuse_from_flow: a bf16 out5 with fp32 frames is inference-only; pass out5.float() "
                           "for a differentiable result")
    if out5.device != img6.device:
        raise RuntimeError("ssm_b200: all tensors of a call must share a device")
    img6, flow4, out5 = _abi.dense_planes(img6.detach()), _abi.dense_planes(flow4.detach()), _abi.dense_planes(out5.detach())
    B, C6, H, W = img6.shape
    if out5.dim() != 5 or C6 != 6 or flow4.shape != (B, 4, H, W) or out5.shape[0] != B or out5.shape[2:] != (5, H, W):
        raise RuntimeError("fuse_from_flow: expected img B x 6, flow B x 4, output B x N x 5 (x H x W), "
                           "got %s, %s, %s" % (tuple(img6.shape), tuple(flow4.shape), tuple(out5.shape)))
    N = out5.shape[1]
    if packed is None and N >= _AUTO_PACK_MIN_TIMESTEPS:
        packed = pack_frames(img6)
    out = _out_buffer(out, (B, N, 3, H, W), img6, "fuse_from_flow")
    with torch.cuda.device(img6.device):
        rc = _abi.lib().ssm_fuse_flow_fwd_mixed(
            _abi.ref(_abi.desc(img6, False)), _packed_ptr(packed, img6), _abi.ref(_abi.desc(flow4, False)),
            _abi.ref(_abi.desc(out5, True)), _abi.DTYPE_BF16, ctypes.c_void_p(tvec.data_ptr()),
            _abi.ref(_abi.desc(out, True)), B, N, H, W, _abi.DTYPE_F32, mode, _abi.stream_ptr(img6.device))
    _abi.check(rc, "ssm_fuse_flow_fwd_mixed")
    return out


def fuse_from_flow(img6, flow4, out5, t, coord_mode=None, packed=None, out=None):
    """Fused frames for every (pair, timestep) from the stage-1 flows: img6 B x 6, flow4 B x 4,
    out5 B x N x 5 -> B x N x 3.  Equals fuse(img6, flow_pack(img6, flow4, t), out5, t) without reading
    the 16-channel tensor back; gradients go to flow4 directly.  A bf16 out5 (the U-Net output under bf16
    autocast) next to fp32 frames and flows is read as it is (inference only)."""
    B, N = out5.shape[0], out5.shape[1]
    if B * N == 0 and img6.is_cuda:
        return _empty_result((B, N, 3) + tuple(img6.shape[2:]), out, img6, flow4, out5)
    tvec = _t_vector(t, B * N, img6.device)
    if out5.dtype == torch.bfloat16 and img6.dtype == torch.float32:
        return _fuse_from_flow_mixed(img6, flow4, out5, tvec, _resolve_mode(coord_mode), packed, out)
    if out is not None:
        _no_grad_for_out(out, "fuse_from_flow", img6, flow4, out5)
        with torch.no_grad():
            return _FuseFlow.forward(_NoCtx(), img6, flow4, out5, tvec, _resolve_mode(coord_mode), packed, out)
    return _FuseFlow.apply(img6, flow4, out5, tvec, _resolve_mode(coord_mode), packed)


# ---------------------------------------------------------------------------------------------
class _FuseLoss(torch.autograd.Function):
    """compute_output_image + the L1 loss front-end of SSMLosses (ssm_fuse_loss_fwd/bwd):
    flow_interpolation.py:394-429 with losses.py:111, 152-167."""

    @staticmethod
    def forward(ctx, img6, flow4, out5, target, tvec, stage1_loss, stage2_loss, mode, packed):
        _same(img6, flow4, out5, target)
        img6, flow4 = _abi.dense_planes(img6), _abi.dense_planes(flow4)
        out5, target = _abi.dense_planes(out5), _abi.dense_planes(target)
        B, C6, H, W = img6.shape
        if out5.dim() != 5 or C6 != 6 or flow4.shape != (B, 4, H, W) or out5.shape[0] != B \
                or out5.shape[2:] != (5, H, W) or target.shape != (B, out5.shape[1], 3, H, W):
            raise RuntimeError("fuse_loss: expected img B x 6, flow B x 4, output B x N x 5, target B x N x 3 "
                               "(x H x W), got %s, %s, %s, %s" % (tuple(img6.shape), tuple(flow4.shape),
                                                                 tuple(out5.shape), tuple(target.shape)))
        N = out5.shape[1]
        if packed is None and N >= _AUTO_PACK_MIN_TIMESTEPS:
            packed = pack_frames(img6)
        L = _abi.lib()
        out = torch.empty((B, N, 3, H, W), dtype=img6.dtype, device=img6.device)
        sums = torch.empty((B, 2 * N + 1), dtype=torch.float32, device=img6.device)
        ws_bytes = L.ssm_fuse_loss_workspace_bytes(B, N, H, W)
        ws = _workspace(ws_bytes, img6.device)
        with torch.cuda.device(img6.device):
            rc = L.ssm_fuse_loss_fwd(_abi.ref(_abi.desc(img6, False)), _packed_ptr(packed, img6),
                                     _abi.ref(_abi.desc(flow4, False)), _abi.ref(_abi.desc(out5, True)),
                                     _abi.ref(_abi.desc(target, True)), ctypes.c_void_p(tvec.data_ptr()),
                                     _abi.ref(_abi.desc(out, True)), ctypes.c_void_p(sums.data_ptr()),
                                     B, N, H, W, _abi.dtype_code(img6), mode, int(stage1_loss), int(stage2_loss),
                                     ctypes.c_void_p(ws.data_ptr()), ws_bytes, _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_fuse_loss_fwd")
        ctx.save_for_backward(img6, flow4, out5, target, tvec, packed, out)
        ctx.mode, ctx.s1, ctx.s2 = mode, int(stage1_loss), int(stage2_loss)
        ctx.set_materialize_grads(False)      # an unused output arrives as None, not as a dense zero tensor
        return out, sums

    @staticmethod
    def backward(ctx, g3, gsums):
        img6, flow4, out5, target, tvec, packed, out = ctx.saved_tensors
        B, _, H, W = img6.shape
        N = out5.shape[1]
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
            raise RuntimeError("fuse_loss treats frames and targets as data (no image gradients): use "
                               "fuse_from_flow + warp for a graph that differentiates the images")
        need_f, need_y = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_f or need_y) or (g3 is None and gsums is None):
            return (None,) * 9
        g3 = _abi.dense_planes(g3.to(img6.dtype)) if g3 is not None else None
        gsums = (gsums if gsums is not None else torch.zeros((B, 2 * N + 1), device=img6.device)).float().contiguous()
        gf = torch.empty_like(flow4, memory_format=torch.contiguous_format) if need_f else None
        gy = torch.empty_like(out5, memory_format=torch.contiguous_format) if need_y else None
        with torch.cuda.device(img6.device):
            rc = _abi.lib().ssm_fuse_loss_bwd(
                _abi.ref(_abi.desc(g3, True)), ctypes.c_void_p(gsums.data_ptr()), _abi.ref(_abi.desc(img6, False)),
                _packed_ptr(packed, img6), _abi.ref(_abi.desc(flow4, False)), _abi.ref(_abi.desc(out5, True)),
                _abi.ref(_abi.desc(target, True)), _abi.ref(_abi.desc(out, True)), ctypes.c_void_p(tvec.data_ptr()),
                _abi.ref(_abi.desc(gy, True)), _abi.ref(_abi.desc(gf, False)), B, N, H, W, _abi.dtype_code(img6),
                ctx.mode, ctx.s1, ctx.s2, _abi.stream_ptr(img6.device))
        _abi.check(rc, "ssm_fuse_loss_bwd")
        return None, gf, gy, None, None, None, None, None, None


def fuse_loss(img6, flow4, out5, target, t, stage1_loss=True, stage2_loss=True, coord_mode=None, packed=None):
    """Fused frames AND the sums of the L1 loss maps of losses.py in one pass.

    img6 B x 6, flow4 B x 4, out5 B x N x 5, target B x N x 3 -> (frames B x N x 3, sums B x (2N+1)) with
    sums[:, 2n] = sum |frame_n - target_n|, sums[:, 2n+1] = stage-2 warp loss sum of timestep n,
    sums[:, 2N] = stage-1 warp loss sum (zero when the corresponding flag is off).  Both outputs are
    differentiable w.r.t. flow4 and out5; frames and targets are data."""
    B, N = out5.shape[0], out5.shape[1]
    tvec = _t_vector(t, B * N, img6.device)
    return _FuseLoss.apply(img6, flow4, out5, target, tvec, bool(stage1_loss), bool(stage2_loss),
                           _resolve_mode(coord_mode), packed)


# ---------------------------------------------------------------------------------------------
def synthesize_host(img6, flow4, out5, t, coord_mode=None, return_inputs=False, out=None, scratch=None):
    """Whole path on HOST tensors (pinned memory recommended) through ssm_synthesize_host: copies
    in, runs the fused kernels for all N timesteps of every pair, copies the frames out.

    img6 B x 6 x H x W, flow4 B x 4 x H x W, out5 B x N x 5 x H x W, t B*N values; fp32 CPU tensors.
    Returns out3 B x N x 3 x H x W (and in16 B x N x 16 x H x W if return_inputs) as CPU tensors.
    out: optional preallocated (pinned) result tensor; scratch: optional uint8 CUDA tensor of
    ssm_synthesize_host_scratch_bytes -- pass both to keep allocations out of repeated calls.
    """
    for name, x in (("img6", img6), ("flow4", flow4), ("out5", out5)):
        if x.device.type != "cpu" or x.dtype != torch.float32 or not x.is_contiguous():
            raise RuntimeError("synthesize_host: %s must be a contiguous fp32 CPU tensor" % name)
    if not torch.cuda.is_available():
        raise RuntimeError("synthesize_host needs a CUDA device -- there is no CPU fallback")
    B, _, H, W = img6.shape
    N = out5.shape[1]
    tv = torch.as_tensor(t, dtype=torch.float32).reshape(-1).contiguous()
    if tv.numel() != B * N:
        raise RuntimeError("synthesize_host: t needs B*N values")
    L = _abi.lib()
    if out is None:
        out = torch.empty((B, N, 3, H, W), dtype=torch.float32, pin_memory=True)
    elif out.shape != (B, N, 3, H, W) or out.dtype != torch.float32 or not out.is_contiguous() or out.is_cuda:
        raise RuntimeError("synthesize_host: out must be a contiguous fp32 CPU tensor of shape B x N x 3 x H x W")
    in16 = torch.empty((B, N, 16, H, W), dtype=torch.float32, pin_memory=True) if return_inputs else None
    need = L.ssm_synthesize_host_scratch_bytes(B, N, H, W)
    if scratch is None:
        scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    elif not scratch.is_cuda or scratch.numel() * scratch.element_size() < need:
        raise RuntimeError("synthesize_host: scratch must be a CUDA tensor of at least %d bytes" % need)
    with torch.cuda.device(scratch.device):
        torch.cuda.current_stream().synchronize()      # scratch may still be in use by earlier work
        rc = L.ssm_synthesize_host(
            ctypes.c_void_p(img6.data_ptr()), ctypes.c_void_p(flow4.data_ptr()), ctypes.c_void_p(out5.data_ptr()),
            ctypes.c_void_p(tv.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(in16.data_ptr()) if in16 is not None else None, B, N, H, W, _resolve_mode(coord_mode),
            ctypes.c_void_p(scratch.data_ptr()), scratch.numel() * scratch.element_size())
    _abi.check(rc, "ssm_synthesize_host")
    return (out, in16) if return_inputs else out


def synthesize_host_scratch_bytes(B, N, H, W):
    return int(_abi.lib().ssm_synthesize_host_scratch_bytes(B, N, H, W))
