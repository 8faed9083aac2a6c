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

    def set_channels_last(self, on=True):
        """SURVEY.md section 8(f) rank 2 (plumbing around the path): run the convolutions in channels-last
        memory format.  The 16-channel input written by compute_inputs is converted once at conv1a; the
        4/5-channel output is converted back to the planar layout the synthesis kernels read."""
        self.channels_last = bool(on)
        self.to(memory_format=torch.channels_last if on else torch.contiguous_format)
        return self

    def _encode(self, x):
        skips = []
        for level in range(1, 6):
            if level > 1:
                x = getattr(self, "pool%d" % level)(x)
            x = getattr(self, "conv%da" % level)(x)
            x = getattr(self, "conv%db" % level)(x)
            skips.append(x)
        return skips, self.pool6(x)

    def _decode(self, h, skips, enc_stage1):
        if self.stage == 2 and self.cross_skip_connect:
            h = torch.cat([h, enc_stage1], dim=1)
        x = self.conv7b(self.conv7a(_up2(h)))
        for level, skip in zip((8, 9, 10, 11), (skips[4], skips[3], skips[2], skips[1])):
            x = _up2(torch.cat([x, skip], dim=1))
            x = getattr(self, "conv%da" % level)(x)
            x = getattr(self, "conv%db" % level)(x)
        x = self.fuse_conv(torch.cat([x, skips[0]], dim=1))
        return self.final_conv(x)

    def forward_flat(self, x, enc_stage1=None, windows=1):
        """x: M x C x H x W (windows/timesteps already folded into M, windows fastest) -> (bottleneck
        M x 512 x H/32 x W/32, output M x C_out x H x W).  `windows` = T: a recurrent bottleneck runs over the
        M/T sequences of T windows."""
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
            if enc_stage1 is not None:
                enc_stage1 = enc_stage1.contiguous(memory_format=torch.channels_last)
        skips, pooled = self._encode(x)
        if isinstance(self.conv6, BiConvRecurrent):
            h = self.conv6(pooled.view(-1, windows, *pooled.shape[1:])).reshape(pooled.shape[0], -1, *pooled.shape[2:])
            if self.channels_last:
                h = h.contiguous(memory_format=torch.channels_last)
        else:
            h = self.conv6(pooled)
        out = self._decode(h, skips, enc_stage1)
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
