"""Timestep-batched forward loop of the full model, with the reference's FullModel call surface
(scripts/models/superslomo_r.py:33-293).

What changes against the reference loop:
  * windows (and, in `interpolate`, timesteps) are folded into the batch: compute_inputs runs once for
    all of them (one launch instead of a Python loop + torch.stack, superslomo_r.py:167-179), and so
    does compute_output_image (:215-238);
  * in inference only the middle window goes through compute_output_image -- the reference computes
    every window and keeps the middle one (:237-238);
  * `interpolate` runs stage 1 ONCE per frame pair for all N intermediate times; the reference's eval
    and visualise loops call the whole model, stage 1 included, once per timestep
    (evaluate_interpolation_results.py:234-242, visualize_interpolation.py:139-144);
  * the inference extras come out of the tensors already at hand instead of being recomputed
    (get_intermediate_outputs, :108-150);
  * compute_output_image takes the stage-1 flows and recomputes input_tensor[:, 6:10] in-kernel
    (ssm_fuse_flow_fwd/bwd) instead of reading the 16-channel tensor back, so in training the flow
    gradient reaches stage 1 without the dense B x 16 gradient (12 zero channels) autograd builds for
    the reference's slice.
The two U-Nets are ordinary torch modules (out of scope); any pair with the reference's interface
can be passed in, e.g. the reference's own FlowComputationModel / FlowInterpolationModel.
"""
import torch
import torch.nn as nn

from . import functional as F_ssm
from . import unets
from .flow_interpolation import SynthesisMixin
from .losses import SSMLosses


def _cfg_bool(cfg, section, key, default):
    if cfg is not None and cfg.has_option(section, key):
        return cfg.getboolean(section, key)
    return default


class FullModel(nn.Module, SynthesisMixin):
    def __init__(self, cfg=None, writer=None, stage1_model=None, stage2_model=None, loss=None):
        super().__init__()
        self.cfg, self.writer = cfg, writer
        self.cross_skip = _cfg_bool(cfg, "STAGE2", "CROSS_SKIP", True)
        load_prev = _cfg_bool(cfg, "STAGE1", "LOADPREV", False)      # superslomo_r.py:46-52
        w1 = cfg.get("STAGE1", "WEIGHTS") if (cfg is not None and load_prev) else None
        w2 = cfg.get("STAGE2", "WEIGHTS") if (cfg is not None and load_prev) else None
        self.stage1_model = stage1_model if stage1_model is not None else unets.get_model(
            w1, 6, 4, self.cross_skip, stage=1, cfg=cfg)
        self.stage2_model = stage2_model if stage2_model is not None else unets.get_model(
            w2, 16, 5, self.cross_skip, stage=2, cfg=cfg)
        self.freeze_weights()
        # The loss module is only needed by training forwards.  The reference builds it eagerly, pretrained VGG16
        # included (losses.py:23); here a VGG16 that cannot be loaded (offline) fails the first TRAINING forward
        # instead of every inference-only construction.
        self._loss_error = None
        if loss is None:
            try:
                loss = SSMLosses(cfg)
            except RuntimeError as e:
                self._loss_error = e
        self.loss = loss
        # interpolate(): write compute_inputs in the layout/dtype a channels-last stage-2 U-Net consumes and read
        # its bf16 output directly; False = planar fp32 either side (generic)
        self.unet_layouts = True

    def freeze_weights(self):
        """superslomo_r.py:73-88"""
        for section, model in (("STAGE1", self.stage1_model), ("STAGE2", self.stage2_model)):
            if _cfg_bool(self.cfg, section, "FREEZE", False):
                model.eval()
                for p in model.parameters():
                    p.requires_grad = False

    @staticmethod
    def get_image_pairs(img_tensor):
        """B x T x 3 x H x W -> B x (T-1) x 6 x H x W, adjacent frames paired (superslomo_r.py:90-106)."""
        return torch.cat([img_tensor[:, :-1], img_tensor[:, 1:]], dim=2)

    # -----------------------------------------------------------------------------------------
    def _stage1(self, image_pairs):
        """-> flows B x W x 4 x H x W, encodings list (or None)."""
        outs = self.stage1_model(image_pairs)
        # the synthesis path runs in the frames' dtype (fp32 unless the caller stores frames in bf16): under
        # autocast the U-Nets return bf16, which would silently cost the coordinates 16 bits
        flows = torch.stack([o[1] for o in outs], dim=1).to(image_pairs.dtype)
        encs = [o[0] for o in outs]
        return flows, encs

    def forward(self, image_tensor, t_interp, target_images=None, iteration=None, inference_mode=True):
        """superslomo_r.py:250-293.  image_tensor B x T x 3 x H x W, t_interp B x (T-1) x 1 x 1 x 1.
        Training: (est_img_t of the middle window, losses [B, 4]); inference: (est_img_t,
        (flowC_01, flowC_10, est_flow_t1, est_flow_t0, refined_flow_t1, refined_flow_t0, v_0t))."""
        if not inference_mode:
            if self.loss is None:
                raise RuntimeError("FullModel: no loss module (%s)" % self._loss_error)
            assert target_images is not None, "No target found for loss."
            assert target_images.shape[1] == image_tensor.shape[1] - 1, "Insufficient number of targets."
        pairs = self.get_image_pairs(image_tensor)                    # B x W x 6 x H x W
        B, Wn = pairs.shape[0], pairs.shape[1]
        mid = Wn // 2
        flows, encs = self._stage1(pairs)
        t_bw = t_interp.reshape(B, Wn)
        flat = lambda x: x.reshape(B * Wn, *x.shape[2:])
        # compute_inputs for every window in one launch (windows folded into the pair axis)
        in16 = F_ssm.flow_pack(flat(pairs), flat(flows), t_bw.reshape(-1), n_timesteps=1)   # (B*W) x 1 x 16
        in16 = in16.view(B, Wn, 16, *in16.shape[-2:])
        out5 = torch.stack(self.stage2_model(in16, encs), dim=1).to(pairs.dtype)      # B x W x 5 x H x W
        if inference_mode:
            sel = slice(mid, mid + 1)
            frame = F_ssm.fuse_from_flow(pairs[:, mid], flows[:, mid], out5[:, sel], t_bw[:, mid])[:, 0]
            x, y = in16[:, mid], out5[:, mid]
            v_0t = 1 - torch.sigmoid(y[:, 0:1])
            extras = (flows[:, mid, 0:2], flows[:, mid, 2:4], x[:, 6:8], x[:, 8:10],
                      x[:, 6:8] + y[:, 1:3], x[:, 8:10] + y[:, 3:5], v_0t)
            return frame, extras
        if isinstance(self.loss, SSMLosses) and not pairs.requires_grad and not target_images.requires_grad:
            # fused: frames + L1 reconstruction + warp losses of every window in one launch
            frames, losses = self.loss.fused_forward(flat(pairs), flat(flows), flat(out5), t_bw.reshape(-1),
                                                     flat(target_images))
            frames = frames.view(B, Wn, 3, *frames.shape[-2:])
            return frames[:, mid], losses.view(B, Wn, 4).sum(dim=1) / Wn
        # the estimated flows are recomputed from the stage-1 flows: no B x W x 16 gradient with 12 zero
        # channels is built for in16, the flow gradient goes to `flows` directly
        frames = F_ssm.fuse_from_flow(flat(pairs), flat(flows), flat(out5).unsqueeze(1), t_bw.reshape(-1))
        frames = frames.view(B, Wn, 3, *frames.shape[-2:])
        losses = 0
        for w in range(Wn):
            losses = losses + self.loss(pairs[:, w], flows[:, w], in16[:, w], out5[:, w], frames[:, w],
                                        target_images[:, w])
        return frames[:, mid], losses / Wn

    # -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def interpolate(self, image_tensor, t_values, unet_chunk=None):
        """All intermediate frames of every sample in one pass: image_tensor B x T x 3 x H x W,
        t_values N times in (0,1) (e.g. k/8, k = 1..7) -> B x N x 3 x H x W (middle window).
        Stage 1 runs once; compute_inputs / compute_output_image take all N times per launch;
        stage 2 runs on the folded (B*N) batch, in chunks of `unet_chunk` timesteps if given."""
        pairs = self.get_image_pairs(image_tensor)
        B, Wn = pairs.shape[0], pairs.shape[1]
        mid = Wn // 2
        t_values = torch.as_tensor(t_values, dtype=torch.float32, device=image_tensor.device).reshape(-1)
        N = t_values.numel()
        flows, encs = self._stage1(pairs)
        t_bn = t_values.view(1, N).expand(B, N).contiguous()
        step = N if unet_chunk is None else max(1, int(unet_chunk))
        rgbx = [F_ssm.pack_frames(pairs[:, w]) for w in range(Wn)]
        # A channels-last stage-2 U-Net gets its input in that layout straight from compute_inputs -- and in
        # bf16 when it runs under bf16 autocast -- so no conversion pass runs in front of conv1a
        # (SURVEY.md section 8(f) rank 2).  One window only: several windows are stacked, which copies anyway.
        nhwc_dtype = None
        if self.unet_layouts and Wn == 1 and getattr(self.stage2_model, "channels_last", False):
            nhwc_dtype = pairs.dtype
            if torch.is_autocast_enabled("cuda"):
                nhwc_dtype = torch.bfloat16 if torch.get_autocast_dtype("cuda") == torch.bfloat16 else None
        frames = []
        for n0 in range(0, N, step):
            tn = t_bn[:, n0:n0 + step].contiguous()
            n = tn.shape[1]
            # B x W x n x 16: every window at every time of the chunk
            if nhwc_dtype is not None:
                in16 = F_ssm.flow_pack_channels_last(pairs[:, 0], flows[:, 0], tn, n_timesteps=n, dtype=nhwc_dtype,
                                                     packed=rgbx[0]).unsqueeze(1)
            elif Wn > 1:
                in16 = torch.stack([F_ssm.flow_pack(pairs[:, w], flows[:, w], tn, n_timesteps=n, packed=rgbx[w])
                                    for w in range(Wn)], dim=1)
            else:
                in16 = F_ssm.flow_pack(pairs[:, 0], flows[:, 0], tn, n_timesteps=n, packed=rgbx[0]).unsqueeze(1)
            # stage 2 sees (B*n) samples of W windows each
            x = in16.permute(0, 2, 1, 3, 4, 5).reshape(B * n, Wn, 16, *in16.shape[-2:])
            e = None
            if encs[0] is not None:
                e = [enc.repeat_interleave(n, dim=0) for enc in encs]
            out5 = self.stage2_model(x, e)[mid]
            # a bf16 U-Net output (autocast) next to fp32 frames is read as it is by the fusion kernel
            if not (out5.dtype == torch.bfloat16 and pairs.dtype == torch.float32):
                out5 = out5.to(pairs.dtype)
            out5 = out5.view(B, n, 5, *in16.shape[-2:])
            frames.append(F_ssm.fuse_from_flow(pairs[:, mid], flows[:, mid], out5, tn, packed=rgbx[mid]))
        return torch.cat(frames, dim=1)

    @torch.no_grad()
    def interpolate_u8(self, images_u8, t_values, order="bgr", unet_chunk=None, as_u8=True, saturate=True):
        """`interpolate` for frames that arrive as 8-bit images, which is how the reference gets them
        (visualize_interpolation.py:61-88 reads them with cv2): images_u8 B x T x H_in x W_in x 3 uint8 on the GPU,
        t_values N times in (0,1) -> the N intermediate frames of the middle window, as B x N x H_in x W_in x 3 uint8
        images in the same channel order (as_u8; de-normalised and cropped as visualize_interpolation.py:221-232,
        264-268) or as B x N x 3 x H x W normalised fp32 frames of the padded size.

        The frames are normalised and padded to a multiple of 32 on the GPU (ssm_frames_from_u8: bit-identical to the
        reference's expression) for the U-Nets and the pass-through channels; the warps of compute_inputs and
        compute_output_image gather from 2 x 2 tables of the raw bytes (ssm_quads_from_u8, one 16-byte request per
        bilinear sample), and the fused frame is written straight as uint8 (ssm_fuse_flow_fwd_q8_u8)."""
        from . import q8
        if images_u8.dim() != 5 or images_u8.shape[-1] != 3 or images_u8.dtype != torch.uint8 or not images_u8.is_cuda:
            raise RuntimeError("interpolate_u8: expected a CUDA uint8 tensor B x T x H x W x 3, got %s %s"
                               % (tuple(images_u8.shape), images_u8.dtype))
        B, T, H_in, W_in, _ = images_u8.shape
        # normalisation table and padding values: built once per device (reading the padding values back synchronises,
        # which must not happen inside a captured or pipelined step)
        cache = self.__dict__.setdefault("_u8_tables", {})
        if images_u8.device not in cache:
            lut = q8.normalisation_lut(device=images_u8.device)
            cache[images_u8.device] = (lut, lut[:, 0].tolist())
        lut, pads = cache[images_u8.device]
        flat = images_u8.reshape(B * T, H_in, W_in, 3)
        planar, _, (top, left) = q8.frames_from_u8(flat, order=order, pad_mode="before", lut=lut, pad_values=pads)
        H, W = planar.shape[-2:]
        norm = q8.norm6()
        pairs = self.get_image_pairs(planar.view(B, T, 3, H, W))
        Wn = T - 1
        mid = Wn // 2
        t_values = torch.as_tensor(t_values, dtype=torch.float32, device=images_u8.device).reshape(-1)
        N = t_values.numel()
        flows, encs = self._stage1(pairs)
        t_bn = t_values.view(1, N).expand(B, N).contiguous()
        step = N if unet_chunk is None else max(1, int(unet_chunk))
        # entry tables of the two frames of every window (for T = 2 the images are already laid out pair by pair)
        quads = [q8.quads_from_u8(images_u8[:, w:w + 2].reshape(B * 2, H_in, W_in, 3).contiguous(), order=order)[0]
                 for w in range(Wn)]
        nhwc_dtype = None
        if self.unet_layouts and Wn == 1 and getattr(self.stage2_model, "channels_last", False):
            nhwc_dtype = torch.float32
            if torch.is_autocast_enabled("cuda"):
                nhwc_dtype = torch.bfloat16 if torch.get_autocast_dtype("cuda") == torch.bfloat16 else None
        out = []
        for n0 in range(0, N, step):
            tn = t_bn[:, n0:n0 + step].contiguous()
            n = tn.shape[1]
            if nhwc_dtype is not None:
                in16 = q8.flow_pack(pairs[:, 0], quads[0], flows[:, 0], tn, norm, n_timesteps=n,
                                    channels_last_dtype=nhwc_dtype).unsqueeze(1)
            elif Wn > 1:
                in16 = torch.stack([q8.flow_pack(pairs[:, w], quads[w], flows[:, w], tn, norm, n_timesteps=n)
                                    for w in range(Wn)], dim=1)
            else:
                in16 = q8.flow_pack(pairs[:, 0], quads[0], flows[:, 0], tn, norm, n_timesteps=n).unsqueeze(1)
            x = in16.permute(0, 2, 1, 3, 4, 5).reshape(B * n, Wn, 16, H, W)
            e = None
            if encs[0] is not None:
                e = [enc.repeat_interleave(n, dim=0) for enc in encs]
            out5 = self.stage2_model(x, e)[mid]
            if out5.dtype != torch.bfloat16:
                out5 = out5.float()
            out5 = out5.view(B, n, 5, H, W)
            if as_u8:
                out.append(q8.fuse_from_flow_to_u8(quads[mid], flows[:, mid], out5, tn, norm, crop=(top, left, H_in, W_in),
                                                   order=order, saturate=saturate))
            else:
                out.append(q8.fuse_from_flow(quads[mid], flows[:, mid], out5, tn, norm))
        return torch.cat(out, dim=1)

