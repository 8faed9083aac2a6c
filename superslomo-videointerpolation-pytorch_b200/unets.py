"""The two flow U-Nets that sit either side of the synthesis path.

OUT OF SCOPE for the B200 rebuild: they are the one dense
contraction of the system and stay on stock PyTorch / cuDNN.  This module only exists so that the
timestep-batched forward loop (superslomo_r.py) can be run, tested and timed on a box where the
reference tree is absent.  Layer names and shapes follow the reference so that its checkpoints
(`stage1_state_dict` / `stage2_state_dict`, scripts/models/unetflow.py:24-30) load unchanged:
  stage 1  scripts/models/flow_computation.py:27-153   6 -> 4 channels (F01, F10)
  stage 2  scripts/models/flow_interpolation.py:27-157 16 -> 5 channels; conv7a takes 1024 channels
           when the stage-1 bottleneck is concatenated in (cross-stage skip, :98-101, :224-228)
The bottleneck is CONV (superslomo_original.ini) or the bidirectional ConvLSTM / ConvGRU of the recurrent
configuration (superslomo_recurrent.ini:97,105; recurrent.py).

Unlike the reference, which loops over the T windows in Python (flow_computation.py:303-325), encoder and
decoder have no coupling between windows, so the T axis is folded into the batch and each layer runs once;
only a recurrent bottleneck sees the windows of a sample as a sequence (B x T x 512 x H/32 x W/32).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import unet_glue
from .layers import avg_pool, conv
from .recurrent import BiConvRecurrent

# (name, in, out, kernel) of the encoder / decoder convolutions
_ENCODER = [("conv1a", None, 32, 7), ("conv1b", 32, 32, 7), ("conv2a", 32, 64, 5), ("conv2b", 64, 64, 5),
            ("conv3a", 64, 128, 3), ("conv3b", 128, 128, 3), ("conv4a", 128, 256, 3), ("conv4b", 256, 256, 3),
            ("conv5a", 256, 512, 3), ("conv5b", 512, 512, 3)]
_DECODER = [("conv7a", None, 512, 3), ("conv7b", 512, 512, 3), ("conv8a", 1024, 256, 3), ("conv8b", 256, 256, 3),
            ("conv9a", 512, 128, 3), ("conv9b", 128, 128, 3), ("conv10a", 256, 64, 3), ("conv10b", 64, 64, 3),
            ("conv11a", 128, 32, 3), ("conv11b", 32, 32, 3)]


def _up2(x):
    """2x bilinear upsampling (flow_computation.py:92-94).  CUDA autocast lists upsample_bilinear2d as an fp32 op:
    left alone it casts a bf16 activation up, interpolates and writes fp32 -- 38 % of a whole 1080p inference step
    (profiles/r01s_pipeline_profile.txt) -- and the next convolution casts the result down again.  The activation
    stays in the dtype the convolutions produced (ATen accumulates the four taps in fp32 either way)."""
    if x.is_cuda and x.dtype != torch.float32 and torch.is_autocast_enabled("cuda"):
        with torch.autocast("cuda", enabled=False):
            return F.interpolate(x, size=(2 * x.shape[2], 2 * x.shape[3]), mode="bilinear", align_corners=False)
    return F.interpolate(x, size=(2 * x.shape[2], 2 * x.shape[3]), mode="bilinear", align_corners=False)


def _has_hooks(m):
    return bool(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or m._backward_pre_hooks)


class FlowUNet(nn.Module):
    """U-Net with five 2x average-pool levels.  forward(unet_in B x T x C x H x W[, stage-1 encodings])
    returns a list of T (encoding-or-None, output) tuples for stage 1 and a list of T outputs for
    stage 2, like the reference models."""

    def __init__(self, in_channels, out_channels, stage, cross_skip=False, conv6=None, bottleneck="CONV"):
        super().__init__()
        assert stage in (1, 2)
        assert bottleneck in ("CONV", "CLSTM", "CGRU"), "BOTTLENECK must be CONV, CLSTM or CGRU"
        self.stage = stage
        self.bottleneck_type = bottleneck
        self.cross_skip_connect = bool(cross_skip)
        for name, cin, cout, k in _ENCODER:
            setattr(self, name, conv(in_channels if cin is None else cin, cout, kernel_size=k, padding=k // 2))
        for i in range(2, 7):
            setattr(self, "pool%d" % i, avg_pool(kernel_size=2))
        if conv6 is not None:
            self.conv6 = conv6
        elif bottleneck == "CONV":
            self.conv6 = nn.Sequential(conv(512, 512, kernel_size=3), conv(512, 512, kernel_size=3))
        else:                                              # flow_computation.py:73-88, flow_interpolation.py:73-88
            self.conv6 = BiConvRecurrent("lstm" if bottleneck == "CLSTM" else "gru", 512, 512, (3, 3), num_layers=2)
        for name, cin, cout, k in _DECODER:
            if cin is None:
                cin = 1024 if (stage == 2 and self.cross_skip_connect) else 512
            setattr(self, name, conv(cin, cout, kernel_size=k))
        self.fuse_conv = conv(64, 32, kernel_size=3)
        self.final_conv = nn.Conv2d(32, out_channels, kernel_size=3, stride=1, padding=1, bias=True)

    channels_last = False     # set_channels_last(): NHWC activations/weights for cuDNN's tensor-core kernels
    # On channels-last CUDA activations the element-wise ops BETWEEN the convolutions (bias add + LeakyReLU, 2x2
    # average pooling, cat + bilinear upsampling) run as this repo's kernels, forward and backward (unet_glue.py);
    # the convolutions stay on cuDNN.  False = stock torch ops everywhere.
    fast_glue = True

    def set_channels_last(self, on=True):
        """SURVEY.md section 8(f) rank 2 (plumbing around the path): run the convolutions in channels-last
        memory format.  The 16-channel input written by compute_inputs is converted once at conv1a; the
        4/5-channel output is converted back to the planar layout the synthesis kernels read."""
        self.channels_last = bool(on)
        self.to(memory_format=torch.channels_last if on else torch.contiguous_format)
        self.invalidate_param_cache()
        return self

    def invalidate_param_cache(self):
        """Drop the cached casts of weights/biases (the bf16 copies the inference fast path keeps).  Called by
        train()/eval(), load_state_dict() and set_channels_last(); call it yourself after writing parameters
        through `.data` (EMA swaps, p.data.copy_()): such writes bump neither the version counter nor data_ptr."""
        self.__dict__["_param_cache"] = {}

    def train(self, mode=True):
        self.invalidate_param_cache()
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_param_cache()
        return super().load_state_dict(*args, **kwargs)

    # ---- fast path: cuDNN convolution without bias + fused bias/LeakyReLU ------------------------------------------
    def _glue_on(self, x):
        return self.fast_glue and self.channels_last and x.is_cuda

    def _params(self, conv_layer, dtype):
        """(weight in the compute dtype, bias as fp32 rounded to the compute dtype -- what autocast hands cuDNN and
        aten::add_), cast once instead of on every call; re-made when the parameters change."""
        w, b = conv_layer.weight, conv_layer.bias

        def make():
            wd = w.detach() if w.dtype == dtype else w.detach().to(dtype)
            return wd.contiguous(memory_format=torch.channels_last), b.detach().to(dtype).float().contiguous()

        if w.dtype == dtype and b.dtype == torch.float32 and w.is_contiguous(memory_format=torch.channels_last):
            # nothing to cast: use the live tensors, so in-place writes through `.data` (EMA swaps, p.data.copy_()),
            # which bump neither the version counter nor data_ptr, are always seen
            return w.detach(), (b.detach() if dtype == torch.float32 else b.detach().to(dtype).float())

        if not isinstance(w, nn.Parameter):
            # an nn.DataParallel replica: its "parameters" are per-forward broadcast copies -- nothing worth keeping,
            # and a cache keyed on them would grow with every call
            return make()
        cache = self.__dict__.setdefault("_param_cache", {})
        key = (id(conv_layer), dtype)
        stamp = (w._version, b._version, w.data_ptr(), b.data_ptr())
        hit = cache.get(key)
        if hit is None or hit[0] != stamp:
            if len(cache) > 256:
                cache.clear()
            hit = (stamp,) + make()
            cache[key] = hit
        return hit[1], hit[2]

    def _block(self, seq, x, into=None):
        """conv + LeakyReLU(0.1) block (layers.conv).  into = (wide, channel_offset, keep): in inference the activation
        is (also) written into that channel slice of the pre-allocated result of a later torch.cat; returns None
        instead of the activation when it was honoured and keep is False."""
        y = self._block_impl(seq, x, into)
        if into is not None and not (isinstance(y, tuple)):
            wide, off, keep = into                     # the fast path did not apply: do what torch.cat would have done
            wide[:, off:off + y.shape[1]] = y
            return y if keep else None
        return y[0] if isinstance(y, tuple) else y

    def _block_impl(self, seq, x, into):
        if not self._glue_on(x) or _has_hooks(seq) or _has_hooks(seq[0]) or _has_hooks(seq[1]):
            return seq(x)             # hooks observe the stock modules: honour them
        dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            return seq(x)
        layer = seq[0]
        recording = torch.is_grad_enabled() and (x.requires_grad or layer.weight.requires_grad or layer.bias.requires_grad)
        if recording:
            # autograd: the parameters themselves enter the graph (autocast casts the weight, the cast is differentiable)
            y = F.conv2d(x, layer.weight, None, layer.stride, layer.padding, layer.dilation, layer.groups)
            if not unet_glue.usable(y):
                return F.leaky_relu_(y + layer.bias.to(y.dtype).view(1, -1, 1, 1), 0.1)
            return unet_glue.bias_leaky_(y, layer.bias, 0.1)
        w, b = self._params(layer, dtype)
        y = F.conv2d(x if x.dtype == dtype else x.to(dtype), w, None, layer.stride, layer.padding, layer.dilation, layer.groups)
        if not unet_glue.usable(y):
            return F.leaky_relu_(y.add_(b.to(dtype).view(1, -1, 1, 1)), 0.1)
        if into is not None and into[0].dtype == y.dtype:
            wide, off, keep = into
            unet_glue.bias_leaky_into(y, b, 0.1, wide, off, keep)
            return (y if keep else None,)              # tuple: `into` has been honoured
        return unet_glue.bias_leaky_(y, b, 0.1)

    def _pool(self, pool, x):
        if self._glue_on(x) and unet_glue.usable(x) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            return unet_glue.avgpool2(x)
        return pool(x)

    def _up2_cat(self, parts):
        """upsample(cat(parts)): flow_computation.py:244-245, 253-254, 262-263, 271-272 upsamples the concatenation of the decoder state and
        the encoder skip; the fast path upsamples each part into its channel slice of the result."""
        if self._glue_on(parts[0]):
            parts = [p if p.dtype == parts[0].dtype else p.to(parts[0].dtype) for p in parts]
            if all(unet_glue.usable(p) for p in parts):
                return unet_glue.upsample2x_cat(parts)
        return _up2(parts[0] if len(parts) == 1 else torch.cat(parts, dim=1))

    def _encode(self, x):
        """-> (skips, pooled bottleneck input, fuse_in).  fuse_in: in inference on the fast path, the pre-allocated input of
        fuse_conv -- cat([conv11b_out, conv1b_out]), flow_computation.py:277 -- whose second half conv1b's activation
        pass fills while it writes the skip; None otherwise (the decoder then calls torch.cat)."""
        skips, fuse_in = [], None
        for level in range(1, 6):
            if level > 1:
                x = self._pool(getattr(self, "pool%d" % level), x)
            x = self._block(getattr(self, "conv%da" % level), x)
            if level == 1 and self._glue_on(x) and not torch.is_grad_enabled() and unet_glue.usable(x):
                fuse_in = torch.empty((x.shape[0], 64, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                                      memory_format=torch.channels_last)
                x = self._block(self.conv1b, x, into=(fuse_in, 32, True))
            else:
                x = self._block(getattr(self, "conv%db" % level), x)
            skips.append(x)
        return skips, self._pool(self.pool6, x), fuse_in

    def _decode(self, h, skips, enc_stage1, fuse_in=None):
        parts = [h, enc_stage1] if (self.stage == 2 and self.cross_skip_connect) else [h]
        x = self._block(self.conv7b, self._block(self.conv7a, self._up2_cat(parts)))
        for level, skip in zip((8, 9, 10, 11), (skips[4], skips[3], skips[2], skips[1])):
            x = self._up2_cat([x, skip])
            x = self._block(getattr(self, "conv%da" % level), x)
            if level == 11 and fuse_in is not None and fuse_in.dtype == x.dtype:
                self._block(self.conv11b, x, into=(fuse_in, 0, False))      # lands in fuse_in[:, 0:32]
                x = None
            else:
                x = self._block(getattr(self, "conv%db" % level), x)
        if x is not None:
            fuse_in = torch.cat([x, skips[0]], dim=1)
        return self.final_conv(self._block(self.fuse_conv, fuse_in))

    def forward_flat(self, x, enc_stage1=None, windows=1):
        """x: M x C x H x W (windows/timesteps already folded into M, windows fastest) -> (bottleneck
        M x 512 x H/32 x W/32, output M x C_out x H x W).  `windows` = T: a recurrent bottleneck runs over the
        M/T sequences of T windows."""
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
            if enc_stage1 is not None:
                enc_stage1 = enc_stage1.contiguous(memory_format=torch.channels_last)
        skips, pooled, fuse_in = self._encode(x)
        if isinstance(self.conv6, BiConvRecurrent):
            h = self.conv6(pooled.view(-1, windows, *pooled.shape[1:])).reshape(pooled.shape[0], -1, *pooled.shape[2:])
            if self.channels_last:
                h = h.contiguous(memory_format=torch.channels_last)
        elif isinstance(self.conv6, nn.Sequential) and len(self.conv6) == 2 and all(isinstance(b, nn.Sequential) for b in self.conv6):
            h = self._block(self.conv6[1], self._block(self.conv6[0], pooled))
        else:
            h = self.conv6(pooled)
        out = self._decode(h, skips, enc_stage1, fuse_in)
        # the synthesis kernels read planar NCHW: hand the (4- or 5-channel) result back in that layout
        return h, (out.contiguous() if self.channels_last else out)

    def forward(self, unet_in, stage1_encoder_output=None):
        B, T = unet_in.shape[0], unet_in.shape[1]
        x = unet_in.reshape(B * T, *unet_in.shape[2:])
        enc = None
        if self.stage == 2 and self.cross_skip_connect:
            enc = torch.stack(list(stage1_encoder_output), dim=1).reshape(B * T, *stage1_encoder_output[0].shape[1:])
        h, out = self.forward_flat(x, enc, windows=T)
        out = out.view(B, T, *out.shape[1:])
        if self.stage == 2:
            return [out[:, w] for w in range(T)]
        h = h.view(B, T, *h.shape[1:])
        return [((h[:, w] if self.cross_skip_connect else None), out[:, w]) for w in range(T)]


def get_model(path, in_channels, out_channels, cross_skip, verbose=False, stage=1, cfg=None):
    """Factory with the reference's signature (scripts/models/unetflow.py:11-32)."""
    section = "STAGE%d" % stage
    bottleneck = "CONV"
    if cfg is not None and cfg.has_option(section, "BOTTLENECK"):
        bottleneck = cfg.get(section, "BOTTLENECK")
    model = FlowUNet(in_channels, out_channels, stage, cross_skip, bottleneck=bottleneck)
    if path is not None:
        data = torch.load(path, map_location="cpu")
        key = "stage%s_state_dict" % stage
        model.load_state_dict(data[key] if key in data else data)
    return model
