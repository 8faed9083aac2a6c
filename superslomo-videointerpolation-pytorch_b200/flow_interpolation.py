"""Drop-in for the per-pixel methods of the reference's FlowInterpolationModel
(scripts/models/flow_interpolation.py:338-429).

`SynthesisMixin` carries `compute_inputs`, `extract_outputs` and `compute_output_image` with the
reference's names, argument order and shapes; mixing it into (or patching it onto) the reference's
FlowInterpolationModel replaces ~110 ATen launches per call by one fused sm_100a kernel each.  The
`*_batched` methods are the timestep-batched forms used by the new forward loop: they take all N
intermediate times of a pair in one launch.
"""
import torch

from . import functional as F_ssm


class SynthesisMixin:
    """Per-pixel synthesis methods; no parameters, no state."""

    def compute_inputs(self, img_tensor, flow_pred_tensor, t):
        """flow_interpolation.py:338.  img_tensor B x 6 x H x W, flow_pred_tensor B x 4 x H x W,
        t B x 1 x 1 x 1 (per-sample time in (0,1)) -> B x 16 x H x W."""
        out = F_ssm.flow_pack(img_tensor, flow_pred_tensor, t, n_timesteps=1)
        return out[:, 0]

    def compute_inputs_batched(self, img_tensor, flow_pred_tensor, t):
        """All N timesteps of every pair at once.  t: B x N -> B x N x 16 x H x W."""
        t = torch.as_tensor(t)
        n = t.shape[1] if t.dim() == 2 else t.numel() // max(img_tensor.shape[0], 1)
        return F_ssm.flow_pack(img_tensor, flow_pred_tensor, t, n_timesteps=n)

    def extract_outputs(self, output_tensor):
        """flow_interpolation.py:374.  Returns (v_1t, dflow_t1, dflow_t0, v_0t).  Kept for callers that
        want the pieces (losses.py:80-101, superslomo_r.py:129-140); compute_output_image does not
        call it -- the sigmoid is fused into the kernel."""
        v_1t = torch.sigmoid(output_tensor[:, 0:1])
        return v_1t, output_tensor[:, 1:3], output_tensor[:, 3:5], 1 - v_1t

    def compute_output_image(self, img_tensor, input_tensor, output_tensor, t):
        """flow_interpolation.py:394.  img_tensor B x 6, input_tensor B x 16, output_tensor B x 5
        (x H x W), t B x 1 x 1 x 1 -> fused frame B x 3 x H x W."""
        # under autocast the stage-2 U-Net hands over a bf16 output_tensor next to fp32 frames; the reference's torch
        # ops promote it (flow_interpolation.py:402-427), so do the same before the single-dtype kernel
        dtype = img_tensor.dtype
        if input_tensor.dtype != dtype:
            input_tensor = input_tensor.to(dtype)
        if output_tensor.dtype != dtype:
            output_tensor = output_tensor.to(dtype)
        out = F_ssm.fuse(img_tensor, input_tensor.unsqueeze(1), output_tensor.unsqueeze(1), t)
        return out[:, 0]

    def compute_output_image_batched(self, img_tensor, input_tensor, output_tensor, t):
        """input_tensor B x N x 16, output_tensor B x N x 5, t B x N -> B x N x 3 x H x W."""
        return F_ssm.fuse(img_tensor, input_tensor, output_tensor, t)

    def compute_output_image_from_flow(self, img_tensor, flow_pred_tensor, output_tensor, t):
        """compute_output_image for callers that still hold flow_pred_tensor (B x 4): the estimated flows
        input_tensor[:, 6:10] are recomputed in-kernel.  output_tensor B x N x 5, t B x N -> B x N x 3."""
        return F_ssm.fuse_from_flow(img_tensor, flow_pred_tensor, output_tensor, t)


def _route(ours, theirs):
    """CUDA tensors take the B200 path; anything else (the reference's CPU evaluation / debugging) keeps running the
    reference's OWN method, untouched -- this package has no CPU implementation of its own."""
    def method(self, first, *args, **kwargs):
        if isinstance(first, torch.Tensor) and first.is_cuda:
            return ours(self, first, *args, **kwargs)
        return theirs(self, first, *args, **kwargs)
    method.__name__ = ours.__name__
    method.__doc__ = ours.__doc__
    method._ssm_b200 = True
    return method


def patch_reference(flow_interpolation_module, layers_module=None, losses_module=None):
    """Install the B200 path into an imported copy of the reference's scripts/models package:

        from models import flow_interpolation, layers, losses
        ssm_b200.patch_reference(flow_interpolation, layers, losses)

    Rebinds FlowInterpolationModel.{compute_inputs, extract_outputs, compute_output_image} and the
    module-level `warp` names the reference calls (flow_interpolation.py:9, layers.py:73, losses.py:8) -- six
    rebindings -- and adds the timestep-batched methods.  CPU tensors still reach the reference's original
    functions (kept under `_ssm_ref_*`), so the reference's CPU evaluation keeps working after patching.
    Idempotent."""
    from .layers import warp as cuda_warp
    cls = flow_interpolation_module.FlowInterpolationModel
    for name in ("compute_inputs", "extract_outputs", "compute_output_image"):
        current = getattr(cls, name)
        if getattr(current, "_ssm_b200", False):
            continue
        setattr(cls, "_ssm_ref_" + name, current)
        setattr(cls, name, _route(getattr(SynthesisMixin, name), current))
    for name in ("compute_inputs_batched", "compute_output_image_batched", "compute_output_image_from_flow"):
        setattr(cls, name, getattr(SynthesisMixin, name))
    for mod in (flow_interpolation_module, layers_module, losses_module):
        if mod is None or getattr(getattr(mod, "warp", None), "_ssm_b200", False):
            continue
        ref_warp = mod.warp

        def warp(x, flo, _ref=ref_warp):
            """layers.warp (scripts/models/layers.py:73): ssm_warp_fwd/bwd for CUDA tensors, the reference's own
            function otherwise"""
            if x.is_cuda:
                return cuda_warp(x, flo)
            return _ref(x, flo)
        warp._ssm_b200 = True
        warp._ssm_ref = ref_warp
        mod.warp = warp
