"""On-disk formats and the frame-directory walk either side of the model loop (SURVEY.md section
8(f) rank 4).  Host-side Python only; nothing here touches the GPU.

  sliding_window     which input frames form each window of a frame directory
                     (reference: scripts/visualize_interpolation.py:270-288)
  output_name        numbering of the frames written (visualize_interpolation.py:226: prefix_%05d.png)
  write_flo/read_flo Middlebury .flo files for the intermediate flows (scripts/utils/flo_utils.py:39-84)
  load_checkpoint    the reference's checkpoint layout: one file with `stage1_state_dict` and
                     `stage2_state_dict` (scripts/main.py:231-237, scripts/models/unetflow.py:24-30)
"""
import os
import struct

import numpy as np
import torch

FLO_MAGIC = 202021.25


def sliding_window(n_images, n_frames, stride=1):
    """Yield, for every adjacent pair (i, i+1) of the (strided) image list, the list of `n_frames` image
    indices centred on that pair, clamped at both ends of the sequence.  stride=8 is the reference's
    240-fps mode (keep every 8th image)."""
    kept = list(range(0, n_images, stride))
    half = (n_frames - 1) // 2
    for start in range(len(kept) - 1):
        locs = [min(max(j, 0), len(kept) - 1) for j in range(start - half, start + 1 + half + 1)]
        yield [kept[j] for j in locs]


def output_name(out_dir, index, prefix="img", ext="png"):
    return os.path.join(out_dir, "%s_%s.%s" % (prefix, str(index).zfill(5), ext))


def write_flo(path, flow):
    """flow: H x W x 2 (u, v) float array or a 2 x H x W / H x W x 2 tensor.  Little-endian, as the reference."""
    if isinstance(flow, torch.Tensor):
        flow = flow.detach().float().cpu()
        if flow.dim() == 3 and flow.shape[0] == 2 and flow.shape[2] != 2:
            flow = flow.permute(1, 2, 0)
        flow = flow.contiguous().numpy()
    flow = np.ascontiguousarray(flow, dtype="<f4")
    h, w, c = flow.shape
    if c != 2:
        raise ValueError("write_flo: expected 2 channels, got %d" % c)
    with open(path, "wb") as f:
        f.write(struct.pack("<f", FLO_MAGIC))
        f.write(struct.pack("<ii", w, h))
        f.write(flow.tobytes())


def read_flo(path):
    with open(path, "rb") as f:
        magic, = struct.unpack("<f", f.read(4))
        if magic != FLO_MAGIC:
            raise ValueError("%s: not a .flo file (magic %r)" % (path, magic))
        w, h = struct.unpack("<ii", f.read(8))
        data = np.frombuffer(f.read(8 * w * h), dtype="<f4")
    return data.reshape(h, w, 2).copy()


def load_checkpoint(model, path, map_location="cpu"):
    """Load a reference checkpoint into a FullModel (or any object with stage1_model / stage2_model)."""
    data = torch.load(path, map_location=map_location)
    model.stage1_model.load_state_dict(data["stage1_state_dict"])
    model.stage2_model.load_state_dict(data["stage2_state_dict"])
    return data.get("iteration", data.get("epoch"))


def save_checkpoint(model, path, **extra):
    d = {"stage1_state_dict": model.stage1_model.state_dict(), "stage2_state_dict": model.stage2_model.state_dict()}
    d.update(extra)
    torch.save(d, path)
