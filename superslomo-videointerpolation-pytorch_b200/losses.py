"""Training losses around the synthesis path, with the reference's call surface
(scripts/models/losses.py:43-249).

Only the warp-loss front-end is part of the rebuilt path: its four `warp` calls
(losses.py:152-161) run through the sm_100a kernel and back-propagate through it.  The L1
reconstruction term is a plain torch expression.  The VGG16 perceptual term (losses.py:12-41, 213) is
torchvision's pretrained VGG16 up to conv4_3, frozen, exactly as the reference builds it; when
LAMBDA_P != 0 and those weights cannot be loaded (no network, no cache) construction FAILS unless the
caller passes a feature extractor or opts out explicitly (`perceptual_features="zero"` or "random").

forward() returns the reference's [B, 4] tensor: (total, reconstruction, warp, perceptual), each the
per-sample mean (kept per sample because the reference gathers it across DataParallel replicas).
"""
import warnings

import torch
import torch.nn as nn

from . import functional as F_ssm
from .layers import warp


def vgg16_conv4_3(pretrained=True):
    """torchvision VGG16 `features[:23]` (up to conv4_3 + ReLU), frozen and in eval mode: the feature extractor of
    the reference's PerceptualLoss (losses.py:12-41).  pretrained=False gives the same architecture with random
    weights (throughput measurements without a network)."""
    import torchvision
    try:
        net = torchvision.models.vgg16(weights=torchvision.models.VGG16_Weights.IMAGENET1K_V1 if pretrained else None)
    except TypeError:          # torchvision < 0.13
        net = torchvision.models.vgg16(pretrained=pretrained)
    feats = net.features[:23].eval()
    for p in feats.parameters():
        p.requires_grad = False
    return feats


def _sample_mean(x):
    return x.reshape(x.shape[0], -1).mean(dim=1, keepdim=True)


class SSMLosses(nn.Module):
    def __init__(self, cfg=None, lambda_r=60.0, lambda_p=20.0, lambda_w=10.0, stage1_frozen=False,
                 stage2_frozen=False, perceptual_features="auto"):
        """perceptual_features: "auto" (default) = the reference's pretrained VGG16 conv4_3 whenever LAMBDA_P != 0,
        raising if the weights are unavailable; a module = use it; "random" = random-init VGG16 conv4_3 (same
        compute, for throughput runs); "zero" / None = explicit opt-out, the perceptual column is zero."""
        super().__init__()
        if cfg is not None:                                    # losses.py:73-78; a partial config keeps the defaults
            get = lambda fn, sec, key, default: fn(sec, key) if cfg.has_option(sec, key) else default
            lambda_r = get(cfg.getfloat, "TRAIN", "LAMBDA_R", lambda_r)
            lambda_w = get(cfg.getfloat, "TRAIN", "LAMBDA_W", lambda_w)
            lambda_p = get(cfg.getfloat, "TRAIN", "LAMBDA_P", lambda_p)
            stage1_frozen = get(cfg.getboolean, "STAGE1", "FREEZE", stage1_frozen)
            stage2_frozen = get(cfg.getboolean, "STAGE2", "FREEZE", stage2_frozen)
        self.loss_weights = (lambda_r, lambda_p, lambda_w)
        self.stage1_frozen, self.stage2_frozen = stage1_frozen, stage2_frozen
        if isinstance(perceptual_features, str):
            kind = perceptual_features.lower()
            if kind not in ("auto", "random", "zero"):
                raise ValueError("perceptual_features must be a module, None, 'auto', 'random' or 'zero'")
            if kind == "zero" or lambda_p == 0:
                perceptual_features = None
            elif kind == "random":
                perceptual_features = vgg16_conv4_3(pretrained=False)
            else:
                try:
                    perceptual_features = vgg16_conv4_3(pretrained=True)
                except Exception as e:      # offline: the reference itself fails here (losses.py:23)
                    raise RuntimeError(
                        "SSMLosses: LAMBDA_P = %g needs torchvision's pretrained VGG16 (reference losses.py:23) and it "
                        "could not be loaded (%s: %s). Pass perceptual_features=<module>, or opt out explicitly with "
                        "perceptual_features='zero' (perceptual term dropped) or 'random' (random-init VGG16)."
                        % (lambda_p, type(e).__name__, str(e)[:120])) from e
        elif perceptual_features is None and lambda_p != 0:
            warnings.warn("SSMLosses: perceptual_features=None with LAMBDA_P = %g: the perceptual term of the reference "
                          "(losses.py:213) is DROPPED and the total loss differs from the reference's" % lambda_p,
                          RuntimeWarning, stacklevel=2)
        self.perceptual_features = perceptual_features     # torchvision vgg16.features[:23], frozen

    def get_warp_loss(self, img_tensor, flowC_output, flowI_input, flowI_output, target_image):
        """losses.py:113-170: (total, stage-1 part, stage-2 part) as B x 3 x H x W L1 maps."""
        img_0, img_1 = img_tensor[:, 0:3], img_tensor[:, 3:6]
        zero = torch.zeros_like(target_image)
        stage1 = stage2 = zero
        if not self.stage1_frozen:
            stage1 = (warp(img_1, flowC_output[:, 0:2]) - img_0).abs() + (warp(img_0, flowC_output[:, 2:4]) - img_1).abs()
        if not self.stage2_frozen:
            flow_t1 = flowI_input[:, 6:8] + flowI_output[:, 1:3]
            flow_t0 = flowI_input[:, 8:10] + flowI_output[:, 3:5]
            stage2 = (warp(img_0, flow_t0) - target_image).abs() + (warp(img_1, flow_t1) - target_image).abs()
        return stage1 + stage2, stage1, stage2

    def forward(self, flowC_input, flowC_output, flowI_input, flowI_output, interpolated_image, target_image):
        lambda_r, lambda_p, lambda_w = self.loss_weights
        rec = _sample_mean(lambda_r * (interpolated_image - target_image).abs())
        if self.perceptual_features is not None and lambda_p != 0:
            fa, fb = self.perceptual_features(interpolated_image), self.perceptual_features(target_image)
            per = _sample_mean(lambda_p * (fa - fb) ** 2)
        else:
            per = torch.zeros_like(rec)
        wl, _, _ = self.get_warp_loss(flowC_input, flowC_output, flowI_input, flowI_output, target_image)
        wl = _sample_mean(lambda_w * wl)
        total = rec + wl + per
        return torch.cat([total, rec, wl, per], dim=1)     # [B, 4]

    def fused_forward(self, img_tensor, flowC_output, flowI_output, t, target_image, packed=None):
        """compute_output_image and this module's forward() in one fused pass (ssm_fuse_loss_fwd/bwd):
        img_tensor B x 6, flowC_output B x 4, flowI_output B x 5, t B values, target_image B x 3
        -> (interpolated_image B x 3, losses [B, 4]).  Same values as
        forward(img, flowC_output, compute_inputs(...), flowI_output, compute_output_image(...), target)
        without the two extra warps (losses.py:152-154) and ~30 elementwise launches per window."""
        lambda_r, lambda_p, lambda_w = self.loss_weights
        frames, sums = F_ssm.fuse_loss(img_tensor, flowC_output, flowI_output.unsqueeze(1), target_image.unsqueeze(1),
                                       t, stage1_loss=not self.stage1_frozen, stage2_loss=not self.stage2_frozen, packed=packed)
        frame = frames[:, 0]
        count = float(target_image[0].numel())
        rec = (lambda_r / count) * sums[:, 0:1]
        wl = (lambda_w / count) * (sums[:, 1:2] + sums[:, 2:3])
        if self.perceptual_features is not None and lambda_p != 0:
            fa, fb = self.perceptual_features(frame), self.perceptual_features(target_image)
            per = _sample_mean(lambda_p * (fa - fb) ** 2)
        else:
            per = torch.zeros_like(rec)
        return frame, torch.cat([rec + wl + per, rec, wl, per], dim=1).to(frame.dtype)
