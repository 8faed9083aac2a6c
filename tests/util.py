"""Shared helpers of the test-suite."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances: max abs error <= 1e-5 in fp32; <= 2e-2 in bf16.  For bf16 STORAGE the bound
# is applied relative to magnitude above 1 (a bf16 value of magnitude m carries a rounding error
# of up to m * 2^-8, which alone exceeds 2e-2 once m > 5).
FP32_TOL = 1e-5
BF16_ATOL = 2e-2
BF16_RTOL = 2.0 ** -7


def golden_cases():
    """hot-path fixtures (warp / compute_inputs / compute_output_image)"""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith(("loop_", "frames_", "q8_", "large_", "sliding_")))


def q8_cases():
    """8-bit-frame fixtures (the reference's load_batch + normalize_tensor + compute_inputs / compute_output_image)"""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "q8_*.npz"))))


def loop_cases():
    """model-loop fixtures (the reference's FullModel, rows a6-a8)"""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if n.startswith("loop_"))


BOTTLENECKS = ("CONV", "CLSTM", "CGRU")     # tests/golden/make_golden.py::BOTTLENECK_CODE


def seeded_unets(seed, device="cpu", bottleneck="CONV"):
    """The two U-Nets with the weights the golden generator's reference FullModel had: same seed,
    same construction order (stage 1 then stage 2), same state_dict keys."""
    from ssm_b200 import unets
    torch.manual_seed(int(seed))
    if not isinstance(bottleneck, str):                      # the code stored in a loop fixture (absent = CONV)
        bottleneck = BOTTLENECKS[int(bottleneck)]
    s1 = unets.FlowUNet(6, 4, 1, cross_skip=True, bottleneck=bottleneck)
    s2 = unets.FlowUNet(16, 5, 2, cross_skip=True, bottleneck=bottleneck)
    return s1.to(device), s2.to(device)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def max_err(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def assert_close_fp32(got, want, what, tol=FP32_TOL):
    e = max_err(got, want)
    assert e <= tol, "%s: max abs error %.3e > %.1e" % (what, e, tol)
    return e


def assert_close_bf16(got, want, what):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    bound = BF16_ATOL + BF16_RTOL * want.abs()
    bad = ((got - want).abs() > bound)
    assert not bad.any(), "%s: %d elements beyond %.0e + 2^-7*|ref| (max abs err %.3e)" % (
        what, int(bad.sum()), BF16_ATOL, (got - want).abs().max().item())
