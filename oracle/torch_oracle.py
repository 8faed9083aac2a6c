"""Torch restatement of the reference's per-pixel synthesis path (ORACLE -- test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.

The reference path is nothing but torch calls, so this restatement issues the same torch
operations in the same order and therefore produces the same bits as the reference on the same
device: on CPU it is the CPU-ATen reference, on a CUDA device with cuDNN disabled it is the
CUDA-ATen reference, with cuDNN enabled the cuDNN spatial-transformer reference (SURVEY.md
findings 3/3b).  It is also what bench.py times as the "reference CPU path" on the GPU box,
where /root/reference does not exist.

Follows:  scripts/models/layers.py:90-120 (warp)
          scripts/models/flow_interpolation.py:349-367 (compute_inputs)
          scripts/models/flow_interpolation.py:382-392, 402-427 (extract_outputs, compute_output_image)
Checked against the imported reference by tests/golden/make_golden.py (bit-equal on CPU).
"""
import torch
import torch.nn.functional as F


def warp(x, flo):
    """layers.py:73-120.  x: B x C x H x W, flo: B x 2 x H x W (channel 0 horizontal)."""
    B, _, H, W = x.shape
    cols = torch.arange(0, W).view(1, 1, 1, W).expand(B, 1, H, W)
    rows = torch.arange(0, H).view(1, 1, H, 1).expand(B, 1, H, W)
    base = torch.cat((cols, rows), 1).float().to(x.device)          # :92-99 (built on the CPU, then moved)
    pos = base + flo                                                # :100
    u = pos[:, 0, :, :].clone()                                     # :109-110
    v = pos[:, 1, :, :].clone()
    u = 2.0 * u / max(W - 1, 1) - 1.0                               # :112
    v = 2.0 * v / max(H - 1, 1) - 1.0                               # :113
    pos[:, 0, :, :] = u                                             # :115-116
    pos[:, 1, :, :] = v
    return F.grid_sample(x, pos.permute(0, 2, 3, 1), align_corners=True)   # :118-119


def compute_inputs(img_tensor, flow_pred_tensor, t):
    """flow_interpolation.py:338-372.  t: B x 1 x 1 x 1."""
    f01 = flow_pred_tensor[:, 0:2]
    f10 = flow_pred_tensor[:, 2:4]
    ft0 = -(1 - t) * t * f01 + (t ** 2) * f10                       # :353
    ft1 = ((1 - t) ** 2) * f01 - t * (1 - t) * f10                  # :356
    i0 = img_tensor[:, 0:3]
    i1 = img_tensor[:, 3:6]
    g1 = warp(i1, ft1)                                              # :361
    g0 = warp(i0, ft0)                                              # :362
    return torch.cat([i1, g1, ft1, ft0, g0, i0], dim=1)             # :364-367


def extract_outputs(output_tensor):
    """flow_interpolation.py:374-392."""
    v1 = torch.sigmoid(output_tensor[:, 0:1])
    return v1, output_tensor[:, 1:3], output_tensor[:, 3:5], 1 - v1


def compute_output_image(img_tensor, input_tensor, output_tensor, t):
    """flow_interpolation.py:394-429."""
    ft1 = input_tensor[:, 6:8]
    ft0 = input_tensor[:, 8:10]
    i0 = img_tensor[:, 0:3]
    i1 = img_tensor[:, 3:6]
    v1, d1, d0, v0 = extract_outputs(output_tensor)
    r1 = ft1 + d1                                                   # :412
    r0 = ft0 + d0                                                   # :413
    p0 = warp(i0, r0)                                               # :416
    p1 = warp(i1, r1)                                               # :418
    p0 = v0 * p0                                                    # :420
    p1 = v1 * p1                                                    # :421
    num = (1 - t) * p0 + t * p1                                     # :423
    den = (1 - t) * v0 + t * v1                                     # :425
    return num / den                                                # :427


# ---- the reference's model loop (scripts/models/superslomo_r.py:90-106, 152-293) ----------------
def get_image_pairs(img_tensor):
    """superslomo_r.py:90-106: B x T x 3 x H x W -> B x (T-1) x 6 x H x W."""
    frames = list(img_tensor.split(dim=1, split_size=1))
    pairs = [torch.cat([a, b], dim=2)[:, 0] for a, b in zip(frames[:-1], frames[1:])]
    return torch.stack(pairs, dim=1)


def warp_loss_maps(img_pair, flow_c, in16, out5, target, stage1_frozen=False, stage2_frozen=False):
    """losses.py:113-170 (L1 maps, before weighting)."""
    i0, i1 = img_pair[:, 0:3], img_pair[:, 3:6]
    total = 0
    if not stage1_frozen:
        total = total + (warp(i1, flow_c[:, 0:2]) - i0).abs() + (warp(i0, flow_c[:, 2:4]) - i1).abs()
    if not stage2_frozen:
        r1 = in16[:, 6:8] + out5[:, 1:3]
        r0 = in16[:, 8:10] + out5[:, 3:5]
        total = total + (warp(i0, r0) - target).abs() + (warp(i1, r1) - target).abs()
    return total


def losses(img_pair, flow_c, in16, out5, frame, target, lambda_r=60.0, lambda_p=20.0, lambda_w=10.0):
    """losses.py:196-249 without the VGG term (weights unavailable offline): [B, 4]."""
    mean = lambda x: x.reshape(x.shape[0], -1).mean(dim=1, keepdim=True)
    rec = mean(lambda_r * (frame - target).abs())
    wl = mean(lambda_w * warp_loss_maps(img_pair, flow_c, in16, out5, target))
    per = torch.zeros_like(rec)
    return torch.cat([rec + wl + per, rec, wl, per], dim=1)


def model_forward(stage1, stage2, image_tensor, t_interp, target_images=None):
    """superslomo_r.py:250-293 + 152-248, window by window as the reference does.  stage1/stage2 are
    modules with the reference's interface.  Returns (middle frame, extras) or (middle frame,
    losses [B, 4]) when targets are given."""
    pairs = get_image_pairs(image_tensor)
    T = pairs.shape[1]
    mid = T // 2
    c_out = stage1(pairs)
    in16, encs = [], []
    for w in range(T):
        enc, flow = c_out[w]
        in16.append(compute_inputs(pairs[:, w], flow, t_interp[:, w]))
        encs.append(enc)
    in16 = torch.stack(in16, dim=1)
    i_out = stage2(in16, encs)
    est, total = None, 0
    for w in range(T):
        frame = compute_output_image(pairs[:, w], in16[:, w], i_out[w], t_interp[:, w])
        if target_images is not None:
            total = total + losses(pairs[:, w], c_out[w][1], in16[:, w], i_out[w], frame, target_images[:, w])
        if w == mid:
            est = frame
    if target_images is not None:
        return est, total / T
    x, y, flow = in16[:, mid], i_out[mid], c_out[mid][1]
    v0 = 1 - torch.sigmoid(y[:, 0:1])
    return est, (flow[:, 0:2], flow[:, 2:4], x[:, 6:8], x[:, 8:10], x[:, 6:8] + y[:, 1:3], x[:, 8:10] + y[:, 3:5], v0)


def interpolate_frames(stage1, stage2, image_tensor, n_intermediate):
    """evaluate_interpolation_results.py:213-244: the whole model once per intermediate time."""
    B, T = image_tensor.shape[0], image_tensor.shape[1]
    out = []
    for idx in range(1, n_intermediate + 1):
        t = torch.full((B, T - 1, 1, 1, 1), float(idx), device=image_tensor.device) / float(n_intermediate + 1)
        out.append(model_forward(stage1, stage2, image_tensor, t)[0])
    return torch.stack(out, dim=1)


# ---- frame pre/post-processing.  Pinned: tests/golden/make_golden_frames.py
# runs the reference's own methods unmodified on CPU (Interpolator.load_batch / normalize_tensor,
# Evaluator.convert_tensor_to_numpy_image, the data loader's Normalize / ToTensor / EvalPad, called unbound
# on a bare object so that the constructors, which build the model on a CUDA device, are not needed),
# asserts that the restatements below are bit-equal and commits tests/golden/frames_prepost.npz.
PIXEL_MEAN = (0.485, 0.456, 0.406)
PIXEL_STD = (0.229, 0.224, 0.225)


def load_batch_and_normalize(bgr_u8, device="cpu"):
    """scripts/visualize_interpolation.py:61-88 (load_batch) then :257-262 (normalize_tensor).
    bgr_u8: T x H x W x 3 uint8 numpy array as cv2.imread returns it -> 1 x T x 3 x H32 x W32 fp32."""
    import numpy as np
    frame_buffer = np.ascontiguousarray(bgr_u8[:, :, :, ::-1])[None, ...]            # :68-72  RGB, 1 T H W C
    frame_buffer = torch.from_numpy(frame_buffer).float().to(device)                 # :73
    frame_buffer = frame_buffer.permute(0, 1, 4, 2, 3)                               # :74
    _, _, _, H, W = frame_buffer.shape
    padding = [0, 0, 0, 0]                                                           # l, r, top, bottom
    if H % 32 != 0:
        h_pad = 32 - (H % 32)
        padding[2] = h_pad // 2
        padding[3] = h_pad - padding[2]
    if W % 32 != 0:
        w_pad = 32 - (W % 32)
        padding[0] = w_pad // 2
        padding[1] = w_pad - padding[0]
    frame_buffer = F.pad(frame_buffer, padding, mode="constant", value=0)            # :87
    pix_mean = torch.tensor(PIXEL_MEAN).view(1, 1, -1, 1, 1).to(device)              # :258-260
    pix_std = torch.tensor(PIXEL_STD).view(1, 1, -1, 1, 1).to(device)
    return (frame_buffer / 255.0 - pix_mean) / pix_std


def reader_normalize_and_pad(rgb_u8, pad_tb):
    """scripts/utils/dataloaders/augmentations.py:181-200 (Normalize on the uint8 numpy array: float64
    arithmetic; ToTensor) then default_reader.py:266-271 (EvalPad(ZeroPad2d([0, 0, pad, pad]))) and the
    .float() the trainer applies.  rgb_u8: T x H x W x 3 -> T x 3 x (H + 2 pad) x W fp32."""
    import numpy as np
    sample = (rgb_u8 / 255.0 - np.asarray(PIXEL_MEAN)) / np.asarray(PIXEL_STD)
    sample = torch.from_numpy(sample.copy()).permute(0, 3, 1, 2)
    return torch.nn.ZeroPad2d([0, 0, pad_tb, pad_tb])(sample).float()


def crop_denormalize_u8(batch, h_start, w_start, h_in, w_in):
    """scripts/evaluate_interpolation_results.py:143-163 (get_crop, convert_tensor_to_numpy_image) and
    :192-202 (denormalize).  batch: B x 3 x H x W fp32 -> B x h_in x w_in x 3 uint8 (RGB)."""
    import numpy as np
    batch = batch.permute(0, 2, 3, 1)
    batch = batch[:, h_start:h_start + h_in, w_start:w_start + w_in, ...]
    pix_mean = torch.tensor(PIXEL_MEAN).view(1, 1, 1, -1).to(batch.device)
    pix_std = torch.tensor(PIXEL_STD).view(1, 1, 1, -1).to(batch.device)
    batch = batch * pix_std + pix_mean
    batch = batch * 255.0
    return batch.cpu().data.numpy().astype(np.uint8)
