"""Element-wise steps between the U-Nets' cuDNN convolutions (ssm_upsample2x_nhwc, ssm_bias_leaky_nhwc,
ssm_avgpool2_nhwc of include/ssm_b200.h), for channels-last activations in inference.

The convolutions stay on PyTorch / cuDNN.  What these replace are the ATen element-wise ops around them, which
take 60 % of a whole 1080p inference step (profiles/r01s_pipeline_profile.txt): the channels-last bilinear
upsampling (the `upsampleN` lambdas of scripts/models/flow_computation.py:92-137, applied at :236-272), the bias add + LeakyReLU(0.1)
of layers.conv (scripts/models/layers.py:21-33) and AvgPool2d(2) (layers.py:60-63).  Same arithmetic, same
operation order as the ATen ops in the forward; the backward kernels are the vector-Jacobian products autograd
derives for those ops, written as gathers (deterministic, no atomics).  `usable(x)` says whether a tensor can take
this path; the callers in unets.py use the stock torch ops otherwise.
"""
import ctypes

import torch

from . import _abi


def usable(x):
    """CUDA, channels-last dense, bf16/fp32, C a multiple of 8."""
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last))


def _ptr(t, elem_offset=0):
    return ctypes.c_void_p(t.data_ptr() + elem_offset * t.element_size())


def _nhwc(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def _nhwc_empty(M, C, H, W, like):
    return torch.empty((M, C, H, W), dtype=like.dtype, device=like.device, memory_format=torch.channels_last)


def _upsample2x_cat_fwd(parts):
    x0 = parts[0]
    M, _, H, W = x0.shape
    C = sum(p.shape[1] for p in parts)
    out = _nhwc_empty(M, C, 2 * H, 2 * W, x0)
    L = _abi.lib()
    off = 0
    with torch.cuda.device_of(x0):
        for p in parts:
            if p.shape[0] != M or p.shape[2:] != x0.shape[2:] or p.dtype != x0.dtype:
                raise RuntimeError("upsample2x_cat: parts must share batch, size and dtype")
            _abi.check(L.ssm_upsample2x_nhwc(_ptr(p), _ptr(out, off), M, H, W, p.shape[1], C, _abi.dtype_code(p),
                                             _abi.stream_ptr(p.device)), "ssm_upsample2x_nhwc")
            off += p.shape[1]
    return out


class _Upsample2xCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *parts):
        ctx.channels = [p.shape[1] for p in parts]
        ctx.size = (parts[0].shape[0], parts[0].shape[2], parts[0].shape[3])
        return _upsample2x_cat_fwd(parts)

    @staticmethod
    def backward(ctx, g):
        g = _nhwc(g)
        M, H, W = ctx.size
        C = sum(ctx.channels)
        L = _abi.lib()
        grads, off = [], 0
        with torch.cuda.device_of(g):
            for i, c in enumerate(ctx.channels):
                if ctx.needs_input_grad[i]:
                    gi = _nhwc_empty(M, c, H, W, g)
                    _abi.check(L.ssm_upsample2x_bwd_nhwc(_ptr(g, off), _ptr(gi), M, H, W, c, C, _abi.dtype_code(g),
                                                         _abi.stream_ptr(g.device)), "ssm_upsample2x_bwd_nhwc")
                    grads.append(gi)
                else:
                    grads.append(None)
                off += c
        return tuple(grads)


def upsample2x_cat(parts):
    """F.interpolate(torch.cat(parts, dim=1), size=(2H, 2W), mode="bilinear", align_corners=False) without the
    concatenated intermediate: every part is upsampled straight into its channel slice of the result (and the
    gradient of every part is gathered straight from its slice of the result's gradient)."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in parts):
        return _Upsample2xCat.apply(*parts)
    return _upsample2x_cat_fwd(parts)


def _bias_leaky_fwd(y, bias_f32, slope):
    M, C, H, W = y.shape
    with torch.cuda.device_of(y):
        _abi.check(_abi.lib().ssm_bias_leaky_nhwc(_ptr(y), _ptr(bias_f32), M * H * W, C, float(slope), _abi.dtype_code(y),
                                                  _abi.stream_ptr(y.device)), "ssm_bias_leaky_nhwc")
    return y


def bias_leaky_into(y, bias_f32, slope, wide, channel_offset, keep_in_place):
    """Inference only: leaky_relu(y + bias) written into channels [channel_offset, channel_offset + C) of the wider
    channels-last tensor `wide` (the result of a torch.cat that then never has to run) -- instead of in place
    (keep_in_place=False) or in addition to the in-place result (keep_in_place=True).  Returns y (valid only with
    keep_in_place) ."""
    M, C, H, W = y.shape
    CW = wide.shape[1]
    if (wide.shape[0], wide.shape[2], wide.shape[3]) != (M, H, W) or wide.dtype != y.dtype or not usable(wide) \
            or channel_offset % 8 or channel_offset + C > CW:
        raise RuntimeError("bias_leaky_into: `wide` must be a channels-last tensor of the same batch, size and dtype")
    with torch.cuda.device_of(y):
        if keep_in_place:
            out1, s1, out2, s2 = _ptr(y), C, _ptr(wide, channel_offset), CW
        else:
            out1, s1, out2, s2 = _ptr(wide, channel_offset), CW, None, 0
        _abi.check(_abi.lib().ssm_bias_leaky_nhwc_to(_ptr(y), _ptr(bias_f32), M * H * W, C, float(slope), out1, s1, out2, s2,
                                                     _abi.dtype_code(y), _abi.stream_ptr(y.device)), "ssm_bias_leaky_nhwc_to")
    return y


class _BiasLeaky(torch.autograd.Function):
    """y <- leaky_relu(y + bias) in place on a convolution output (which the convolution's own backward does not
    need); saves the OUTPUT, whose sign is the pre-activation's."""

    @staticmethod
    def forward(ctx, y, bias, slope):
        # the bias joins in the activation's dtype, as autocast hands it to aten::add_
        _bias_leaky_fwd(y, bias.detach().to(y.dtype).float().contiguous(), slope)
        ctx.mark_dirty(y)
        ctx.save_for_backward(y)
        ctx.slope, ctx.bias_dtype = slope, bias.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gy = _nhwc(gy)
        M, C, H, W = y.shape
        gx = torch.empty_like(y, memory_format=torch.channels_last)
        with torch.cuda.device_of(y):
            _abi.check(_abi.lib().ssm_leaky_bwd_nhwc(_ptr(gy), _ptr(y), _ptr(gx), M * H * W, C, float(ctx.slope),
                                                     _abi.dtype_code(y), _abi.stream_ptr(y.device)), "ssm_leaky_bwd_nhwc")
        gb = gx.sum(dim=(0, 2, 3), dtype=torch.float32).to(ctx.bias_dtype) if ctx.needs_input_grad[1] else None
        return gx, gb, None


def bias_leaky_(y, bias, slope=0.1):
    """y <- leaky_relu(y + bias, slope) in place.  Inference: `bias` is an fp32 CUDA tensor of C values already rounded
    to the activation dtype.  With autograd recording: `bias` is the convolution's bias parameter."""
    if torch.is_grad_enabled() and (y.requires_grad or bias.requires_grad):
        return _BiasLeaky.apply(y, bias, slope)
    return _bias_leaky_fwd(y, bias, slope)


def _avgpool2_fwd(x):
    M, C, H, W = x.shape
    out = _nhwc_empty(M, C, H // 2, W // 2, x)
    with torch.cuda.device_of(x):
        _abi.check(_abi.lib().ssm_avgpool2_nhwc(_ptr(x), _ptr(out), M, H // 2, W // 2, C, _abi.dtype_code(x),
                                                _abi.stream_ptr(x.device)), "ssm_avgpool2_nhwc")
    return out


class _AvgPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return _avgpool2_fwd(x)

    @staticmethod
    def backward(ctx, g):
        g = _nhwc(g)
        M, C, H, W = ctx.shape
        gin = _nhwc_empty(M, C, H, W, g)
        with torch.cuda.device_of(g):
            _abi.check(_abi.lib().ssm_avgpool2_bwd_nhwc(_ptr(g), _ptr(gin), M, H // 2, W // 2, C, _abi.dtype_code(g),
                                                        _abi.stream_ptr(g.device)), "ssm_avgpool2_bwd_nhwc")
        return gin


def avgpool2(x):
    """AvgPool2d(2) of a channels-last tensor with even H and W."""
    if torch.is_grad_enabled() and x.requires_grad:
        return _AvgPool2.apply(x)
    return _avgpool2_fwd(x)


# ---- drop-in for the reference's own U-Net instances -------------------------------------------------------------
class FusedConvLeaky(torch.nn.Sequential):
    """nn.Sequential(Conv2d, LeakyReLU) -- what layers.conv builds (scripts/models/layers.py:21-33) -- with the same
    children and therefore the same state_dict keys ("0.weight", "0.bias"), whose forward runs the convolution on
    cuDNN WITHOUT its bias and adds bias + LeakyReLU in one in-place pass.  Falls back to the stock children whenever
    the activation cannot take that pass (CPU, planar layout, other dtypes, C % 8 != 0, hooks installed)."""

    def forward(self, x):
        conv, act = self[0], self[1]
        if (not x.is_cuda or conv.bias is None or conv._forward_hooks or conv._forward_pre_hooks
                or act._forward_hooks or act._forward_pre_hooks):
            return act(conv(x))
        y = torch.nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if not usable(y):
            return act(y + conv.bias.to(y.dtype).view(1, -1, 1, 1))
        if torch.is_grad_enabled() and (y.requires_grad or conv.bias.requires_grad):
            return bias_leaky_(y, conv.bias, act.negative_slope)
        return bias_leaky_(y, conv.bias.detach().to(y.dtype).float().contiguous(), act.negative_slope)


class FastAvgPool2(torch.nn.AvgPool2d):
    """AvgPool2d(2) (layers.avg_pool, scripts/models/layers.py:60-63) on the channels-last kernel when it applies."""

    def forward(self, x):
        if usable(x) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and not (self._forward_hooks or self._forward_pre_hooks):
            return avgpool2(x)
        return super().forward(x)


def _fast_upsample(x):
    if usable(x):
        return upsample2x_cat([x])
    if x.is_cuda and x.dtype != torch.float32 and torch.is_autocast_enabled("cuda"):
        with torch.autocast("cuda", enabled=False):      # autocast would interpolate in fp32 (see unets._up2)
            return torch.nn.functional.interpolate(x, size=(2 * x.shape[2], 2 * x.shape[3]), mode="bilinear", align_corners=False)
    return torch.nn.functional.interpolate(x, size=(2 * x.shape[2], 2 * x.shape[3]), mode="bilinear", align_corners=False)


def accelerate_unet(model):
    """Apply the element-wise kernels to an instance of the REFERENCE's FlowComputationModel / FlowInterpolationModel
    (scripts/models/flow_computation.py, flow_interpolation.py) in place: parameters, state_dict keys and results
    (up to rounding) are unchanged, every convolution stays an nn.Conv2d on cuDNN.
      * every nn.Sequential(Conv2d, LeakyReLU) child becomes a FusedConvLeaky over the same two modules,
      * every AvgPool2d(2) child becomes a FastAvgPool2,
      * the `upsample7` .. `upsample11` lambdas (flow_computation.py:92-137) become the 2x bilinear kernel,
      * weights are converted to channels-last, which makes cuDNN produce channels-last activations.
    Returns the model."""
    nn = torch.nn
    for name, child in list(model.named_children()):
        if (isinstance(child, nn.Sequential) and not isinstance(child, FusedConvLeaky) and len(child) == 2
                and isinstance(child[0], nn.Conv2d) and isinstance(child[1], nn.LeakyReLU)):
            setattr(model, name, FusedConvLeaky(child[0], child[1]))
        elif isinstance(child, nn.Sequential):
            accelerate_unet(child)                      # e.g. the CONV bottleneck: Sequential(conv(...), conv(...))
        elif (isinstance(child, nn.AvgPool2d) and not isinstance(child, FastAvgPool2) and child.kernel_size in (2, (2, 2))
              and child.stride in (2, (2, 2), None) and child.padding in (0, (0, 0)) and not child.ceil_mode):
            setattr(model, name, FastAvgPool2(2))
    for level in range(7, 12):
        if hasattr(model, "upsample%d" % level) and not isinstance(getattr(model, "upsample%d" % level), nn.Module):
            setattr(model, "upsample%d" % level, _fast_upsample)
    model.to(memory_format=torch.channels_last)
    return model
