"""ssm_b200 -- B200 (sm_100a) implementation of Super SloMo's per-pixel intermediate-frame
synthesis path behind the reference's own call surface (scripts/models/layers.py,
scripts/models/flow_interpolation.py:338-429, scripts/models/superslomo_r.py).  CUDA only; the
kernels live in libssm_b200.so (C ABI: include/ssm_b200.h)."""
from . import _abi
from .functional import (flow_pack, flow_pack_channels_last, fuse, fuse_from_flow, fuse_loss, get_coord_mode, pack_frames, pack_image, set_coord_mode,
                         set_device_t_check, synthesize_host, synthesize_host_scratch_bytes, t_violations)
from .layers import avg_pool, conv, warp
from .flow_interpolation import SynthesisMixin, patch_reference
from . import formats, frames, losses, q8, sharding, superslomo_r, synthetic, unet_glue, unets
from .unet_glue import accelerate_unet
from .frames import frames_from_u8, frames_to_u8, normalisation_lut
from .superslomo_r import FullModel

__all__ = ["warp", "conv", "avg_pool", "flow_pack", "flow_pack_channels_last", "fuse", "fuse_from_flow", "fuse_loss", "pack_frames", "pack_image", "synthesize_host",
           "synthesize_host_scratch_bytes", "SynthesisMixin", "FullModel", "patch_reference",
           "set_coord_mode", "get_coord_mode", "set_device_t_check", "t_violations", "abi_version", "frames_from_u8", "frames_to_u8",
           "normalisation_lut", "accelerate_unet"]


def abi_version():
    return int(_abi.lib().ssm_version())
