"""ctypes binding of libssm_b200.so (the C ABI declared in include/ssm_b200.h).

There is no fallback of any kind: if the shared library is missing, or a call fails, a
RuntimeError is raised.  PyTorch supplies device memory and streams only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSM_B200_LIB points at an alternative build of the same library (kernel-variant experiments)
LIB_PATH = os.environ.get("SSM_B200_LIB") or os.path.join(_HERE, "libssm_b200.so")

DTYPE_F32, DTYPE_BF16 = 0, 1
COORD_DIV, COORD_RCP = 0, 1

# every symbol include/ssm_b200.h declares
EXPORTS = (
    "ssm_version", "ssm_last_error",
    "ssm_warp_fwd", "ssm_warp_bwd", "ssm_packed_image_bytes", "ssm_pack_image", "ssm_warp_fwd_packed", "ssm_warp_bwd_packed",
    "ssm_flow_pack_fwd", "ssm_flow_pack_bwd", "ssm_flow_pack_fwd_nhwc",
    "ssm_fuse_fwd", "ssm_fuse_bwd", "ssm_fuse_flow_fwd", "ssm_fuse_flow_bwd", "ssm_fuse_flow_fwd_mixed",
    "ssm_fuse_loss_fwd", "ssm_fuse_loss_bwd", "ssm_fuse_loss_workspace_bytes",
    "ssm_frames_from_u8", "ssm_frames_to_u8",
    "ssm_warp_bwd_workspace_bytes", "ssm_flow_pack_bwd_workspace_bytes", "ssm_fuse_bwd_workspace_bytes",
    "ssm_packed_frames_bytes", "ssm_pack_frames",
    "ssm_synthesize_host", "ssm_synthesize_host_scratch_bytes", "ssm_selftest_division",
    "ssm_upsample2x_nhwc", "ssm_bias_leaky_nhwc", "ssm_avgpool2_nhwc",
    "ssm_upsample2x_bwd_nhwc", "ssm_leaky_bwd_nhwc", "ssm_avgpool2_bwd_nhwc", "ssm_bias_leaky_nhwc_to",
    "ssm_quads_bytes", "ssm_quads_from_u8", "ssm_flow_pack_fwd_q8", "ssm_flow_pack_fwd_q8_nhwc", "ssm_flow_pack_fwd_q8_lut", "ssm_fuse_flow_fwd_q8",
    "ssm_fuse_flow_fwd_q8_u8", "ssm_synthesize_host_u8", "ssm_synthesize_host_u8_scratch_bytes",
)


class SsmTensor(ctypes.Structure):
    """struct ssm_tensor of include/ssm_b200.h"""
    _fields_ = [("data", ctypes.c_void_p), ("stride_b", ctypes.c_int64),
                ("stride_n", ctypes.c_int64), ("stride_c", ctypes.c_int64)]


_lib = None


def lib():
    """Load libssm_b200.so (once).  Raises if it has not been built -- no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libssm_b200.so is missing at %s. Build it with `python __graft_entry__.py` or "
            "`python superslomo-videointerpolation-pytorch_b200/build.py`; this package has no "
            "CPU or eager fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    P, I, V, Z = ctypes.POINTER(SsmTensor), ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
    L.ssm_version.restype = I
    L.ssm_last_error.restype = ctypes.c_char_p
    L.ssm_warp_fwd.argtypes = [P, P, P, I, I, I, I, I, I, V]
    L.ssm_warp_bwd.argtypes = [P, P, P, P, P, I, I, I, I, I, I, V, Z, V]
    L.ssm_packed_image_bytes.argtypes = [I, I, I, I]
    L.ssm_packed_image_bytes.restype = Z
    L.ssm_pack_image.argtypes = [P, V, I, I, I, I, V]
    L.ssm_warp_fwd_packed.argtypes = [V, P, P, I, I, I, I, I, V]
    L.ssm_warp_bwd_packed.argtypes = [P, V, P, P, P, I, I, I, I, I, V, Z, V]
    L.ssm_selftest_division.argtypes = [I, V, V]
    L.ssm_selftest_division.restype = I
    L.ssm_pack_frames.argtypes = [P, V, I, I, I, I, V]
    L.ssm_flow_pack_fwd.argtypes = [P, V, P, V, P, I, I, I, I, I, I, V]
    L.ssm_flow_pack_fwd_nhwc.argtypes = [P, V, P, V, V, I, I, I, I, I, I, I, V]
    L.ssm_flow_pack_bwd.argtypes = [P, P, V, P, V, P, P, I, I, I, I, I, I, V, Z, V]
    L.ssm_fuse_fwd.argtypes = [P, V, P, P, V, P, I, I, I, I, I, I, V]
    L.ssm_fuse_bwd.argtypes = [P, P, V, P, P, V, P, P, P, I, I, I, I, I, I, V, Z, V]
    L.ssm_fuse_flow_fwd.argtypes = L.ssm_fuse_fwd.argtypes
    L.ssm_fuse_flow_bwd.argtypes = L.ssm_fuse_bwd.argtypes
    L.ssm_fuse_flow_fwd_mixed.argtypes = [P, V, P, P, I, V, P, I, I, I, I, I, I, V]
    L.ssm_fuse_loss_fwd.argtypes = [P, V, P, P, P, V, P, V, I, I, I, I, I, I, I, I, V, Z, V]
    L.ssm_fuse_loss_bwd.argtypes = [P, V, P, V, P, P, P, P, V, P, P, I, I, I, I, I, I, I, I, V]
    F3, LL = ctypes.POINTER(ctypes.c_float), ctypes.c_longlong
    L.ssm_frames_from_u8.argtypes = [V, LL, I, I, I, I, I, I, I, I, I, V, F3, P, V, I, V]
    L.ssm_frames_to_u8.argtypes = [P, I, I, I, I, I, I, I, F3, F3, ctypes.c_float, I, I, V, LL, I, I, V]
    for n in ("ssm_warp_bwd_workspace_bytes", "ssm_flow_pack_bwd_workspace_bytes", "ssm_fuse_bwd_workspace_bytes",
              "ssm_packed_frames_bytes", "ssm_synthesize_host_scratch_bytes", "ssm_fuse_loss_workspace_bytes"):
        getattr(L, n).argtypes = [I, I, I, I]
        getattr(L, n).restype = Z
    L.ssm_synthesize_host.argtypes = [V, V, V, V, V, V, I, I, I, I, I, V, Z]
    L.ssm_quads_bytes.argtypes = [I, I, I]
    L.ssm_quads_bytes.restype = Z
    L.ssm_quads_from_u8.argtypes = [V, LL, I, I, I, I, I, I, I, I, I, V, V]
    L.ssm_flow_pack_fwd_q8.argtypes = [P, V, P, V, P, F3, I, I, I, I, I, I, V]
    L.ssm_flow_pack_fwd_q8_nhwc.argtypes = [P, V, P, V, V, I, F3, I, I, I, I, I, I, V]
    L.ssm_flow_pack_fwd_q8_lut.argtypes = [V, V, P, V, P, F3, I, I, I, I, I, I, V]
    L.ssm_fuse_flow_fwd_q8.argtypes = [V, P, P, I, V, P, F3, I, I, I, I, I, I, V]
    L.ssm_fuse_flow_fwd_q8_u8.argtypes = [V, P, P, I, V, V, LL, I, I, I, I, I, F3, F3, ctypes.c_float, I, I, F3, I, I, I, I, I, I, V]
    L.ssm_synthesize_host_u8_scratch_bytes.argtypes = [I, I, I, I, I, I, I]
    L.ssm_synthesize_host_u8_scratch_bytes.restype = Z
    L.ssm_synthesize_host_u8.argtypes = [V, I, V, V, I, V, V, F3, F3, F3, F3, I, I, I, I, I, I, I, I, I, I, V, Z]
    L.ssm_upsample2x_nhwc.argtypes = [V, V, I, I, I, I, LL, I, V]
    L.ssm_bias_leaky_nhwc.argtypes = [V, V, LL, I, ctypes.c_float, I, V]
    L.ssm_avgpool2_nhwc.argtypes = [V, V, I, I, I, I, I, V]
    L.ssm_bias_leaky_nhwc_to.argtypes = [V, V, LL, I, ctypes.c_float, V, LL, V, LL, I, V]
    L.ssm_upsample2x_bwd_nhwc.argtypes = [V, V, I, I, I, I, LL, I, V]
    L.ssm_leaky_bwd_nhwc.argtypes = [V, V, V, LL, I, ctypes.c_float, I, V]
    L.ssm_avgpool2_bwd_nhwc.argtypes = [V, V, I, I, I, I, I, V]
    for n in ("ssm_warp_fwd", "ssm_warp_bwd", "ssm_pack_image", "ssm_warp_fwd_packed", "ssm_warp_bwd_packed", "ssm_pack_frames", "ssm_flow_pack_fwd", "ssm_flow_pack_bwd",
              "ssm_flow_pack_fwd_nhwc", "ssm_fuse_flow_fwd_mixed",
              "ssm_fuse_fwd", "ssm_fuse_bwd", "ssm_fuse_flow_fwd", "ssm_fuse_flow_bwd", "ssm_fuse_loss_fwd", "ssm_fuse_loss_bwd",
              "ssm_frames_from_u8", "ssm_frames_to_u8", "ssm_synthesize_host",
              "ssm_upsample2x_nhwc", "ssm_bias_leaky_nhwc", "ssm_avgpool2_nhwc",
              "ssm_upsample2x_bwd_nhwc", "ssm_leaky_bwd_nhwc", "ssm_avgpool2_bwd_nhwc", "ssm_bias_leaky_nhwc_to",
              "ssm_quads_from_u8", "ssm_flow_pack_fwd_q8", "ssm_flow_pack_fwd_q8_nhwc", "ssm_flow_pack_fwd_q8_lut", "ssm_fuse_flow_fwd_q8",
              "ssm_fuse_flow_fwd_q8_u8", "ssm_synthesize_host_u8"):
        getattr(L, n).restype = I
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().ssm_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def dtype_code(t):
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise TypeError("ssm_b200 supports float32 and bfloat16 storage, got %s" % t.dtype)


def desc(t, has_n):
    """ssm_tensor for a (B, C, H, W) [has_n=False] or (B, N, C, H, W) [has_n=True] CUDA tensor whose
    H x W planes are dense.  Returns None for t is None (an unwanted gradient)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ssm_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
    st = t.stride()
    W = t.shape[-1]
    if st[-1] != 1 or st[-2] != W:
        raise RuntimeError("ssm_b200 needs dense H x W planes (strides %s for shape %s)" % (st, tuple(t.shape)))
    if has_n:
        return SsmTensor(t.data_ptr(), st[0], st[1], st[2])
    return SsmTensor(t.data_ptr(), st[0], 0, st[1])


def ref(d):
    return ctypes.byref(d) if d is not None else None


def dense_planes(t):
    """Return t itself if its planes are dense, else a contiguous copy."""
    st = t.stride()
    if st[-1] == 1 and st[-2] == t.shape[-1]:
        return t
    return t.contiguous()


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
