"""The synthesis path for frames that arrive as 8-bit images (include/ssm_b200.h, "8-bit frames").

The reference reads uint8 images and normalises them on the fly (scripts/visualize_interpolation.py:61-88,
257-262).  That normalisation is affine in the byte, so it commutes with bilinear interpolation: the kernels
behind this module gather RAW BYTES from a table of 2x2 entries (one 16-byte request per bilinear sample
instead of four) and normalise after interpolating.  Same results as flow_pack / fuse_from_flow on the
normalised fp32 frames to within 1e-6; inference only (training crops go through the fp32 kernels, whose
backward exists).

    planar, quads, norm, (top, left) = q8.prepare(images_u8)          # F x H_in x W_in x 3 uint8 CUDA tensor
    in16 = q8.flow_pack(planar.view(B, 6, H, W), quads, flow4, t, norm, n_timesteps=N)
    frames = q8.fuse_from_flow(quads, flow4, out5, t, norm)            # fp32 B x N x 3 x H x W
    images = q8.fuse_from_flow_to_u8(quads, flow4, out5, t, norm, crop=(top, left, H_in, W_in))
"""
import ctypes

import numpy as np
import torch

from . import _abi
from .frames import center_padding, frames_from_u8, normalisation_lut
from .functional import _resolve_mode, _t_vector
from .synthetic import PIXEL_MEAN, PIXEL_STD


def norm6(mean=PIXEL_MEAN, std=PIXEL_STD, divisor=255.0):
    """The affine form of the reference's normalisation, (b / divisor - mean) / std = a * b + c, as the six host
    floats {a_R, a_G, a_B, c_R, c_G, c_B} the q8 entry points take (computed in float64, rounded once)."""
    m, s = np.asarray(mean, dtype=np.float64), np.asarray(std, dtype=np.float64)
    return (ctypes.c_float * 6)(*[float(v) for v in np.concatenate([1.0 / (divisor * s), -m / s])])


def _f3(values):
    return (ctypes.c_float * 3)(*[float(v) for v in values])


def quads_from_u8(images, order="bgr", multiple=32):
    """images: F x H_in x W_in x 3 uint8 CUDA tensor -> (entry tables F x (H+1) x (W+1) x 16 uint8, (H, W, top, left)).
    The images are centred in an H x W frame (multiples of `multiple`, visualize_interpolation.py:76-85) whose
    other pixels are byte 0 (padding BEFORE normalising, :87 then :137)."""
    if not images.is_cuda or images.dtype != torch.uint8:
        raise RuntimeError("quads_from_u8 needs a uint8 CUDA tensor (no CPU fallback); got %s on %s" % (images.dtype, images.device))
    if images.dim() != 4 or images.shape[-1] != 3 or images.stride(-1) != 1 or images.stride(-2) != 3:
        raise RuntimeError("quads_from_u8: expected F x H x W x 3 with dense pixels, got %s / strides %s"
                           % (tuple(images.shape), images.stride()))
    F, h_in, w_in, _ = images.shape
    H, W, top, left = center_padding(h_in, w_in, multiple)
    quads = torch.empty((F, H + 1, W + 1, 16), dtype=torch.uint8, device=images.device)
    with torch.cuda.device(images.device):
        rc = _abi.lib().ssm_quads_from_u8(ctypes.c_void_p(images.data_ptr()), images.stride(0), images.stride(1),
                                          1 if order.lower() == "bgr" else 0, F, h_in, w_in, H, W, top, left,
                                          ctypes.c_void_p(quads.data_ptr()), _abi.stream_ptr(images.device))
    _abi.check(rc, "ssm_quads_from_u8")
    return quads, (H, W, top, left)


def prepare(images, order="bgr", lut=None, mean=PIXEL_MEAN, std=PIXEL_STD, multiple=32, pad_values=None):
    """Everything the q8 kernels need from F uint8 images: (planar normalised frames F x 3 x H x W fp32,
    entry tables, norm6, (top, left)).  The planar frames feed the stage-1 U-Net and the pass-through channels
    of compute_inputs; they are bit-identical to the reference's normalisation (ssm_frames_from_u8)."""
    if lut is None:
        lut = normalisation_lut(mean=mean, std=std, device=images.device)
    planar, _, (top, left) = frames_from_u8(images, order=order, pad_mode="before", lut=lut, multiple=multiple,
                                            pad_values=pad_values)
    quads, _ = quads_from_u8(images, order=order, multiple=multiple)
    return planar, quads, norm6(mean, std), (top, left)


def _check_quads(quads, B, H, W, device):
    if quads.dtype != torch.uint8 or not quads.is_cuda or quads.device != device or not quads.is_contiguous() \
            or quads.numel() != B * 2 * (H + 1) * (W + 1) * 16:
        raise RuntimeError("q8: quads must be the contiguous uint8 entry tables of %d frame pairs at %d x %d "
                           "(%d bytes) on %s" % (B, H, W, B * 2 * (H + 1) * (W + 1) * 16, device))
    return ctypes.c_void_p(quads.data_ptr())


def _flow_pack_from_tables(quads, lut, flow4, t, norm, n_timesteps, coord_mode, out):
    """flow_pack with img6=None: the pass-through channels are looked up from the tables' own bytes (ssm_flow_pack_fwd_q8_lut)"""
    if not flow4.is_cuda or flow4.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("q8.flow_pack: flow4 must be a CUDA tensor, fp32 or bf16 (no CPU fallback)")
    if torch.is_grad_enabled() and flow4.requires_grad:
        raise RuntimeError("q8.flow_pack is inference-only; use flow_pack on fp32 frames for a differentiable result")
    if not (isinstance(lut, torch.Tensor) and lut.is_cuda and lut.device == flow4.device and lut.dtype == torch.float32
            and tuple(lut.shape) == (3, 256) and lut.is_contiguous()):
        raise RuntimeError("q8.flow_pack: lut must be the contiguous 3 x 256 fp32 normalisation table on %s "
                           "(ssm_b200.normalisation_lut)" % flow4.device)
    flow4 = _abi.dense_planes(flow4.detach())
    B, C4, H, W = flow4.shape
    N = int(n_timesteps)
    if C4 != 4:
        raise RuntimeError("compute_inputs: expected flow B x 4 x H x W, got %s" % (tuple(flow4.shape),))
    qp = _check_quads(quads, B, H, W, flow4.device)
    tvec = _t_vector(t, B * N, flow4.device)
    if out is None:
        out = torch.empty((B, N, 16, H, W), dtype=flow4.dtype, device=flow4.device)
    elif tuple(out.shape) != (B, N, 16, H, W) or out.dtype != flow4.dtype or not out.is_contiguous():
        raise RuntimeError("q8.flow_pack: out= must be a contiguous %s %s tensor" % (flow4.dtype, (B, N, 16, H, W)))
    with torch.cuda.device(flow4.device):
        rc = _abi.lib().ssm_flow_pack_fwd_q8_lut(qp, ctypes.c_void_p(lut.data_ptr()), _abi.ref(_abi.desc(flow4, False)),
                                                 ctypes.c_void_p(tvec.data_ptr()), _abi.ref(_abi.desc(out, True)), norm,
                                                 B, N, H, W, _abi.dtype_code(flow4), _resolve_mode(coord_mode),
                                                 _abi.stream_ptr(flow4.device))
    _abi.check(rc, "ssm_flow_pack_fwd_q8_lut")
    return out


def flow_pack(img6, quads, flow4, t, norm, n_timesteps=1, coord_mode=None, out=None, channels_last_dtype=None, lut=None):
    """compute_inputs (flow_interpolation.py:338-372) for n_timesteps times of every pair, warping through the
    entry tables: img6 B x 6 x H x W fp32 (normalised frames; read for the pass-through channels only),
    quads = tables of the same 2B frames, flow4 B x 4 x H x W -> B x N x 16 x H x W (fp32, or bf16 storage throughout).
    channels_last_dtype (torch.float32 / torch.bfloat16): write B x N x H x W x 16 in that dtype instead (returned as a
    B x N x 16 x H x W view), the layout a channels-last stage-2 U-Net consumes.
    img6=None with lut= (the 3 x 256 table of normalisation_lut that made the frames): the pass-through channels are
    looked up from the tables' own bytes and no planar frame is read -- the same values when the frames were padded with
    byte 0 before normalising (prepare(..., pad_values=lut[:, 0])), 5 % less DRAM traffic (planar layout only)."""
    if img6 is None:
        if lut is None or channels_last_dtype is not None:
            raise RuntimeError("q8.flow_pack: img6=None needs lut= and the planar layout")
        return _flow_pack_from_tables(quads, lut, flow4, t, norm, n_timesteps, coord_mode, out)
    if not (img6.is_cuda and flow4.is_cuda) or img6.dtype not in (torch.float32, torch.bfloat16) or flow4.dtype != img6.dtype:
        raise RuntimeError("q8.flow_pack: img6 and flow4 must be CUDA tensors of one storage dtype, fp32 or bf16 (no CPU fallback)")
    if torch.is_grad_enabled() and (img6.requires_grad or flow4.requires_grad):
        raise RuntimeError("q8.flow_pack is inference-only; use flow_pack on fp32 frames for a differentiable result")
    img6, flow4 = _abi.dense_planes(img6.detach()), _abi.dense_planes(flow4.detach())
    B, C6, H, W = img6.shape
    N = int(n_timesteps)
    if C6 != 6 or flow4.shape != (B, 4, H, W):
        raise RuntimeError("compute_inputs: expected img B x 6 x H x W and flow B x 4 x H x W, got %s and %s"
                           % (tuple(img6.shape), tuple(flow4.shape)))
    qp = _check_quads(quads, B, H, W, img6.device)
    tvec = _t_vector(t, B * N, img6.device)
    L = _abi.lib()
    with torch.cuda.device(img6.device):
        if channels_last_dtype is None:
            if out is None:
                out = torch.empty((B, N, 16, H, W), dtype=img6.dtype, device=img6.device)
            elif tuple(out.shape) != (B, N, 16, H, W) or out.dtype != img6.dtype or not out.is_contiguous():
                raise RuntimeError("q8.flow_pack: out= must be a contiguous %s %s tensor" % (img6.dtype, (B, N, 16, H, W)))
            rc = L.ssm_flow_pack_fwd_q8(_abi.ref(_abi.desc(img6, False)), qp, _abi.ref(_abi.desc(flow4, False)),
                                        ctypes.c_void_p(tvec.data_ptr()), _abi.ref(_abi.desc(out, True)), norm,
                                        B, N, H, W, _abi.dtype_code(img6), _resolve_mode(coord_mode), _abi.stream_ptr(img6.device))
        else:
            if channels_last_dtype not in (torch.float32, torch.bfloat16):
                raise TypeError("q8.flow_pack: channels_last_dtype must be float32 or bfloat16")
            if out is None:
                out = torch.empty((B, N, H, W, 16), dtype=channels_last_dtype, device=img6.device).permute(0, 1, 4, 2, 3)
            elif tuple(out.shape) != (B, N, 16, H, W) or out.dtype != channels_last_dtype \
                    or out.stride() != (N * 16 * H * W, 16 * H * W, 1, 16 * W, 16):
                raise RuntimeError("q8.flow_pack: out= must be a %s %s tensor stored as B x N x H x W x 16"
                                   % ((B, N, 16, H, W), channels_last_dtype))
            rc = L.ssm_flow_pack_fwd_q8_nhwc(_abi.ref(_abi.desc(img6, False)), qp, _abi.ref(_abi.desc(flow4, False)),
                                             ctypes.c_void_p(tvec.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                             _abi.DTYPE_BF16 if channels_last_dtype == torch.bfloat16 else _abi.DTYPE_F32,
                                             norm, B, N, H, W, _abi.dtype_code(img6), _resolve_mode(coord_mode),
                                             _abi.stream_ptr(img6.device))
    _abi.check(rc, "ssm_flow_pack_fwd_q8")
    return out


def _fuse_args(quads, flow4, out5, t):
    if not (flow4.is_cuda and out5.is_cuda) or flow4.dtype not in (torch.float32, torch.bfloat16) \
            or out5.dtype not in (torch.float32, torch.bfloat16) or (flow4.dtype == torch.bfloat16 and out5.dtype != torch.bfloat16):
        raise RuntimeError("q8.fuse_from_flow: flow4 fp32 with out5 fp32 or bf16, or both bf16, on CUDA (no CPU fallback)")
    if torch.is_grad_enabled() and (flow4.requires_grad or out5.requires_grad):
        raise RuntimeError("q8.fuse_from_flow is inference-only; use fuse_from_flow on fp32 frames for a differentiable result")
    flow4, out5 = _abi.dense_planes(flow4.detach()), _abi.dense_planes(out5.detach())
    B, _, H, W = flow4.shape
    if out5.dim() != 5 or flow4.shape[1] != 4 or out5.shape[0] != B or out5.shape[2:] != (5, H, W):
        raise RuntimeError("fuse_from_flow: expected flow B x 4 and output B x N x 5 (x H x W), got %s, %s"
                           % (tuple(flow4.shape), tuple(out5.shape)))
    N = out5.shape[1]
    return flow4, out5, B, N, H, W, _check_quads(quads, B, H, W, flow4.device), _t_vector(t, B * N, flow4.device)


def fuse_from_flow(quads, flow4, out5, t, norm, coord_mode=None, out=None):
    """extract_outputs + compute_output_image (flow_interpolation.py:374-429) for every (pair, timestep) through the
    entry tables: flow4 B x 4 fp32, out5 B x N x 5 (fp32, or bf16 as autocast leaves it) -> B x N x 3 x H x W fp32."""
    flow4, out5, B, N, H, W, qp, tvec = _fuse_args(quads, flow4, out5, t)
    if out is None:
        out = torch.empty((B, N, 3, H, W), dtype=flow4.dtype, device=flow4.device)
    elif tuple(out.shape) != (B, N, 3, H, W) or out.dtype != flow4.dtype or not out.is_contiguous():
        raise RuntimeError("q8.fuse_from_flow: out= must be a contiguous %s %s tensor" % (flow4.dtype, (B, N, 3, H, W)))
    with torch.cuda.device(flow4.device):
        rc = _abi.lib().ssm_fuse_flow_fwd_q8(qp, _abi.ref(_abi.desc(flow4, False)), _abi.ref(_abi.desc(out5, True)),
                                             _abi.dtype_code(out5), ctypes.c_void_p(tvec.data_ptr()),
                                             _abi.ref(_abi.desc(out, True)), norm, B, N, H, W, _abi.dtype_code(flow4),
                                             _resolve_mode(coord_mode), _abi.stream_ptr(flow4.device))
    _abi.check(rc, "ssm_fuse_flow_fwd_q8")
    return out


def fuse_from_flow_to_u8(quads, flow4, out5, t, norm, crop=None, mean=PIXEL_MEAN, std=PIXEL_STD, scale=255.0,
                         order="bgr", saturate=True, coord_mode=None, out=None):
    """fuse_from_flow with frames_to_u8 fused behind it: -> B x N x h_out x w_out x 3 uint8 images (crop = (top, left,
    h_out, w_out), default the whole frame), de-normalised as visualize_interpolation.py:221-232, 264-268."""
    flow4, out5, B, N, H, W, qp, tvec = _fuse_args(quads, flow4, out5, t)
    top, left, h_out, w_out = (0, 0, H, W) if crop is None else crop
    if out is None:
        out = torch.empty((B, N, h_out, w_out, 3), dtype=torch.uint8, device=flow4.device)
    elif tuple(out.shape) != (B, N, h_out, w_out, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise RuntimeError("q8.fuse_from_flow_to_u8: out= must be a contiguous uint8 %s tensor" % ((B, N, h_out, w_out, 3),))
    with torch.cuda.device(flow4.device):
        rc = _abi.lib().ssm_fuse_flow_fwd_q8_u8(
            qp, _abi.ref(_abi.desc(flow4, False)), _abi.ref(_abi.desc(out5, True)), _abi.dtype_code(out5),
            ctypes.c_void_p(tvec.data_ptr()), ctypes.c_void_p(out.data_ptr()), h_out * w_out * 3, w_out * 3, top, left,
            h_out, w_out, _f3(mean), _f3(std), float(scale), 1 if order.lower() == "bgr" else 0, 1 if saturate else 0,
            norm, B, N, H, W, _abi.dtype_code(flow4), _resolve_mode(coord_mode), _abi.stream_ptr(flow4.device))
    _abi.check(rc, "ssm_fuse_flow_fwd_q8_u8")
    return out


def synthesize_host(frames_u8, flow4, out5, t, order="bgr", mean=PIXEL_MEAN, std=PIXEL_STD, saturate=True,
                    coord_mode=None, out=None, scratch=None, multiple=32):
    """Whole path on HOST tensors for 8-bit frames (ssm_synthesize_host_u8): frames_u8 B x 2 x H_in x W_in x 3 uint8,
    flow4 B x 4 x H x W fp32 and out5 B x N x 5 x H x W (fp32 or bf16) at the padded size, t B*N values
    -> B x N x H_in x W_in x 3 uint8 interpolated images.  Pinned memory recommended."""
    for name, x, dts in (("frames_u8", frames_u8, (torch.uint8,)), ("flow4", flow4, (torch.float32,)),
                         ("out5", out5, (torch.float32, torch.bfloat16))):
        if x.device.type != "cpu" or x.dtype not in dts or not x.is_contiguous():
            raise RuntimeError("q8.synthesize_host: %s must be a contiguous CPU tensor of dtype %s" % (name, dts))
    if not torch.cuda.is_available():
        raise RuntimeError("q8.synthesize_host needs a CUDA device -- there is no CPU fallback")
    B, two, h_in, w_in, _ = frames_u8.shape
    H, W, top, left = center_padding(h_in, w_in, multiple)
    N = out5.shape[1]
    if two != 2 or flow4.shape != (B, 4, H, W) or out5.shape != (B, N, 5, H, W):
        raise RuntimeError("q8.synthesize_host: expected frames B x 2 x h x w x 3, flow4 B x 4 x %d x %d, out5 B x N x 5 x %d x %d"
                           % (H, W, H, W))
    tv = torch.as_tensor(t, dtype=torch.float32).reshape(-1).contiguous()
    if tv.numel() != B * N:
        raise RuntimeError("q8.synthesize_host: t needs B*N values")
    L = _abi.lib()
    if out is None:
        out = torch.empty((B, N, h_in, w_in, 3), dtype=torch.uint8, pin_memory=True)
    elif tuple(out.shape) != (B, N, h_in, w_in, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or out.is_cuda:
        raise RuntimeError("q8.synthesize_host: out must be a contiguous uint8 CPU tensor of shape B x N x h x w x 3")
    code5 = _abi.DTYPE_BF16 if out5.dtype == torch.bfloat16 else _abi.DTYPE_F32
    need = L.ssm_synthesize_host_u8_scratch_bytes(B, N, h_in, w_in, H, W, code5)
    if scratch is None:
        scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    elif not scratch.is_cuda or scratch.numel() * scratch.element_size() < need:
        raise RuntimeError("q8.synthesize_host: scratch must be a CUDA tensor of at least %d bytes" % need)
    lut = normalisation_lut(mean=mean, std=std, device="cpu").contiguous()
    with torch.cuda.device(scratch.device):
        torch.cuda.current_stream().synchronize()
        rc = L.ssm_synthesize_host_u8(
            ctypes.c_void_p(frames_u8.data_ptr()), 1 if order.lower() == "bgr" else 0, ctypes.c_void_p(flow4.data_ptr()),
            ctypes.c_void_p(out5.data_ptr()), code5, ctypes.c_void_p(tv.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            ctypes.cast(lut.data_ptr(), ctypes.POINTER(ctypes.c_float)), norm6(mean, std), _f3(mean), _f3(std),
            1 if saturate else 0, B, N, h_in, w_in, H, W, top, left, _resolve_mode(coord_mode),
            ctypes.c_void_p(scratch.data_ptr()), scratch.numel() * scratch.element_size())
    _abi.check(rc, "ssm_synthesize_host_u8")
    return out


def synthesize_host_scratch_bytes(B, N, h_in, w_in, out5_dtype=torch.float32, multiple=32):
    H, W, _, _ = center_padding(h_in, w_in, multiple)
    return int(_abi.lib().ssm_synthesize_host_u8_scratch_bytes(
        B, N, h_in, w_in, H, W, _abi.DTYPE_BF16 if out5_dtype == torch.bfloat16 else _abi.DTYPE_F32))
