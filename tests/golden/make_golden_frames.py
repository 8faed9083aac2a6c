"""Generates tests/golden/frames_prepost.npz by running the UNMODIFIED reference pre/post steps on CPU.

Run in the build container only (the reference lives at /root/reference and does not travel):

    python tests/golden/make_golden_frames.py

The reference keeps these steps inside classes whose constructors build the full model on a CUDA device
(Interpolator, Evaluator) and whose modules import packages that are absent here.  The script therefore
  * registers empty stand-ins for the absent, unrelated imports (matplotlib, skimage, more_itertools,
    tensorboardX) so that the reference modules import,
  * makes `Tensor.cuda()` the identity for the duration of the run (there is no GPU here), and
  * calls the reference METHODS unbound on a bare object carrying only the attributes they read --
    the method bodies that run are the reference's own bytes:
      Interpolator.load_batch / normalize_tensor / denormalize_tensor   scripts/visualize_interpolation.py:61-88, 257-268
      Evaluator.convert_tensor_to_numpy_image (get_crop + denormalize)   scripts/evaluate_interpolation_results.py:143-163, 192-202
      augmentations.Normalize / ToTensor                                scripts/utils/dataloaders/augmentations.py:181-200
      default_reader.EvalPad                                            scripts/utils/dataloaders/default_reader.py
Inputs are seeded uint8 images written as PNG (lossless) and read back by the reference's own cv2.imread.
It also checks, before writing, that the restatements in oracle/torch_oracle.py reproduce the reference
bit for bit; tests/test_oracle_golden.py and tests/test_frames_gpu.py then use the fixture as the pin.
"""
import configparser
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/scripts"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


# absent packages the reference modules import at the top but the pre/post steps never touch
_stub("matplotlib"); _stub("matplotlib.pyplot"); _stub("matplotlib.colors")
_stub("skimage"); _stub("skimage.measure", compare_psnr=None, compare_ssim=None)
_stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
_stub("more_itertools", windowed=None)
_stub("tensorboardX", SummaryWriter=object)

torch.Tensor.cuda = lambda self, *a, **k: self          # no GPU in the build container


def _import_reference():
    import cv2  # noqa: F401  (the reference reads the images with it)
    mods = {}
    for name in ("visualize_interpolation", "evaluate_interpolation_results",
                 "utils.dataloaders.augmentations", "utils.dataloaders.default_reader"):
        try:
            mods[name] = __import__(name, fromlist=["x"])
        except Exception as e:  # say exactly what blocks the import instead of guessing
            raise SystemExit("cannot import reference module %s: %s: %s" % (name, type(e).__name__, e))
    return mods


def main():
    import cv2
    mods = _import_reference()
    vis, ev = mods["visualize_interpolation"], mods["evaluate_interpolation_results"]
    aug, reader = mods["utils.dataloaders.augmentations"], mods["utils.dataloaders.default_reader"]
    from oracle import torch_oracle

    rng = np.random.RandomState(1234)
    out = {}

    # ---- Interpolator.load_batch + normalize_tensor (visualise path), ragged size 45 x 70 -> 64 x 96 ----
    T, h_in, w_in = 3, 45, 70
    bgr = rng.randint(0, 256, size=(T, h_in, w_in, 3)).astype(np.uint8)
    bgr[0, 0, 0] = (0, 128, 255)
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i in range(T):
            p = os.path.join(d, "%05d.png" % i)
            assert cv2.imwrite(p, bgr[i])
            paths.append(p)
        bare = types.SimpleNamespace()
        loaded = vis.Interpolator.load_batch(bare, paths)                      # 1 T 3 H32 W32, 0..255
    normalised = vis.Interpolator.normalize_tensor(bare, loaded)
    out["vis_bgr_u8"] = bgr
    out["vis_normalised"] = normalised.numpy()
    mine = torch_oracle.load_batch_and_normalize(bgr)
    assert mine.shape == normalised.shape and torch.equal(mine, normalised), "load_batch_and_normalize differs"

    # ---- Interpolator.denormalize_tensor --------------------------------------------------------------
    x5 = torch.randn(1, 2, 3, 32, 64, generator=torch.Generator().manual_seed(7)) * 1.2
    den = vis.Interpolator.denormalize_tensor(bare, x5)
    out["vis_denorm_in"] = x5.numpy()
    out["vis_denorm_out"] = den.numpy()

    # ---- Evaluator.convert_tensor_to_numpy_image: crop + denormalise + astype(uint8) --------------------
    cfg = configparser.RawConfigParser()
    cfg.read("/root/reference/configs/superslomo_original.ini")
    # H_REF = ceil(H_IN / 32) * 32, H_START = (H_REF - H_IN) // 2 (evaluate_interpolation_results.py:86-93)
    bare_ev = types.SimpleNamespace(cfg=cfg, H_START=9, W_START=13, H_IN=45, W_IN=70, H_REF=64, W_REF=96)
    bare_ev.get_crop = types.MethodType(ev.Evaluator.get_crop, bare_ev)
    bare_ev.denormalize = types.MethodType(ev.Evaluator.denormalize, bare_ev)
    xb = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(8)) * 1.3
    u8 = ev.Evaluator.convert_tensor_to_numpy_image(bare_ev, xb)
    out["eval_in"] = xb.numpy()
    out["eval_crop"] = np.asarray([9, 13, 45, 70])
    out["eval_u8"] = u8
    mine = torch_oracle.crop_denormalize_u8(xb, 9, 13, 45, 70)
    assert mine.shape == u8.shape and np.array_equal(mine, u8), "crop_denormalize_u8 differs"

    # ---- data-loader path: augmentations.Normalize + ToTensor, then EvalPad -----------------------------
    rgb = rng.randint(0, 256, size=(2, 40, 64, 3)).astype(np.uint8)
    pix_mean = [float(p) for p in cfg.get("MODEL", "PIXEL_MEAN").split(",")]
    pix_std = [float(p) for p in cfg.get("MODEL", "PIXEL_STD").split(",")]
    sample = aug.Normalize(pix_mean, pix_std)(rgb)
    sample = aug.ToTensor()(sample)
    pad = 12
    padded = reader.EvalPad(torch.nn.ZeroPad2d([0, 0, pad, pad]))(sample).float()
    out["reader_rgb_u8"] = rgb
    out["reader_pad"] = np.asarray([pad])
    out["reader_out"] = padded.numpy()
    mine = torch_oracle.reader_normalize_and_pad(rgb, pad)
    assert mine.shape == padded.shape and torch.equal(mine, padded), "reader_normalize_and_pad differs"

    path = os.path.join(HERE, "frames_prepost.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "oracle/torch_oracle.py pre/post restatements are bit-equal to the reference on CPU")


if __name__ == "__main__":
    main()
