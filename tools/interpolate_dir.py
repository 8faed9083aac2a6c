#!/usr/bin/env python
"""Interpolate a directory of frames (what the reference's scripts/visualize_interpolation.py does):

    python tools/interpolate_dir.py --input-dir frames/ --output-dir out/ --upsample-rate 8 [--weights ckpt.pt]
                                    [--n-frames 2] [--fps-240] [--save-flows] [--amp] [--channels-last] [--bottleneck CLSTM]

uint8 images go to the GPU as they are; FullModel.interpolate_u8 normalises and pads them for the U-Nets
(ssm_frames_from_u8), runs stage 1 once per window and all intermediate times per launch with the warps gathering
the raw bytes (ssm_quads_from_u8, ssm_flow_pack_fwd_q8), and writes the fused frames straight as cropped uint8 images
(ssm_fuse_flow_fwd_q8_u8).  Without --weights the U-Nets are random-init (seed 42): useful only as a smoke run.
"""
import argparse
import glob
import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssm_b200  # noqa: E402
from ssm_b200 import formats  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input-dir", required=True)
    ap.add_argument("--output-dir", required=True)
    ap.add_argument("--img-type", default="png")
    ap.add_argument("--upsample-rate", type=int, default=8)
    ap.add_argument("--n-frames", type=int, default=2)
    ap.add_argument("--fps-240", action="store_true")
    ap.add_argument("--weights")
    ap.add_argument("--save-flows", action="store_true")
    ap.add_argument("--amp", action="store_true")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--bottleneck", default="CONV", choices=["CONV", "CLSTM", "CGRU"],
                    help="CONV: superslomo_original.ini; CLSTM: superslomo_recurrent.ini (use --n-frames 4)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    import configparser
    cfg = configparser.RawConfigParser()
    cfg.read_string("[STAGE1]\nBOTTLENECK=%s\n[STAGE2]\nBOTTLENECK=%s\nCROSS_SKIP=TRUE\n" % (a.bottleneck, a.bottleneck))
    model = ssm_b200.FullModel(cfg=cfg, loss=ssm_b200.losses.SSMLosses(cfg, perceptual_features="zero")).to(dev).eval()   # inference only
    if a.weights:
        formats.load_checkpoint(model, a.weights)
    if a.channels_last:
        model.stage1_model.set_channels_last()
        model.stage2_model.set_channels_last()
    paths = sorted(glob.glob(os.path.join(a.input_dir, "*." + a.img_type.lower())))
    os.makedirs(a.output_dir, exist_ok=True)
    t_values = [k / a.upsample_rate for k in range(1, a.upsample_rate)]
    count = 0
    last = None
    for window in formats.sliding_window(len(paths), a.n_frames, stride=8 if a.fps_240 else 1):
        imgs = np.stack([cv2.imread(paths[i]) for i in window])                     # T x H x W x 3, BGR
        T, H, W, _ = imgs.shape
        images = torch.from_numpy(imgs).to(dev)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            out = model.interpolate_u8(images[None], t_values, order="bgr")[0].cpu().numpy()   # N x H x W x 3 uint8, BGR
        mid = T // 2 - 1
        cv2.imwrite(formats.output_name(a.output_dir, count), imgs[mid]); count += 1
        for k in range(out.shape[0]):
            cv2.imwrite(formats.output_name(a.output_dir, count), out[k]); count += 1
        last = imgs[mid + 1]
        if a.save_flows:
            planar, _, (top, left) = ssm_b200.frames_from_u8(images, order="bgr")
            with torch.no_grad():
                flows, _ = model._stage1(model.get_image_pairs(planar[None]))
            f = flows[0, T // 2 - 1]
            formats.write_flo(os.path.join(a.output_dir, "flow_01_%05d.flo" % count), f[0:2, top:top + H, left:left + W])
            formats.write_flo(os.path.join(a.output_dir, "flow_10_%05d.flo" % count), f[2:4, top:top + H, left:left + W])
    if last is not None:
        cv2.imwrite(formats.output_name(a.output_dir, count), last)
    print("wrote %d frames to %s" % (count + 1, a.output_dir))


if __name__ == "__main__":
    main()
