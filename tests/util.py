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


def load_large_golden(name="large_352_flow100"):
    """the large-flow fixture (tests/golden/make_golden_large.py): exact fp32 inputs rebuilt from the stored integers
    with the generator's own expressions, outputs for the stored band of rows"""
    d = load_golden(name)
    mean = torch.tensor((0.485, 0.456, 0.406)).view(1, 3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(1, 3, 1, 1)
    x = (d["u8"].permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    d["img6"] = torch.cat([x[0::2], x[1::2]], dim=1).contiguous()
    d["flow4"] = d["flow_i16"].float() / 64.0
    d["out5"] = d["out5_i16"].float() / 1024.0
    r0, r1 = (int(v) for v in d["band"])
    B, _, H, W = d["flow4"].shape
    d["g16"] = torch.zeros((B, 16, H, W))
    d["g16"][:, 3:13, r0:r1] = d["g16_band_i16"].float() / 256.0
    d["g3"] = torch.zeros((B, 3, H, W))
    d["g3"][:, :, r0:r1] = d["g3_band_i16"].float() / 256.0
    d["rows"] = slice(r0, r1)
    return d


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


def assert_sum_close(got, terms, what, tol=FP32_TOL):
    """`got` is a SUM of the fp32 tensors `terms` that the kernel accumulated in registers (e.g. the flow gradient
    over N timesteps).  The yardstick is the same fp32 terms added in float64, which is exact for this purpose; any
    fp32 accumulation of them -- the kernel's, or the one autograd performs for the reference -- may differ from it by
    up to len(terms) roundings of the running sum, i.e. len(terms) * 2^-24 * sum|terms| per element.  The bar is
    north_star's 1e-5 plus that representation bound (printed next to the error of a plain fp32 left-to-right sum of
    the same terms, the reference's own accumulation noise)."""
    terms = [x.detach().cpu() for x in terms]
    ref64 = sum(x.double() for x in terms)
    mag = sum(x.double().abs() for x in terms)
    bound = len(terms) * 2.0 ** -24 * mag
    fp32_sum = terms[0].clone()
    for x in terms[1:]:
        fp32_sum = fp32_sum + x
    noise = (fp32_sum.double() - ref64).abs().max().item()
    err = (got.detach().cpu().double() - ref64).abs()
    bad = err > tol + bound
    assert not bad.any(), "%s: %d elements beyond %.0e + %d ulp-roundings of the running sum (max err %.3e, fp32 reference sum noise %.3e)" % (
        what, int(bad.sum()), tol, len(terms), err.max().item(), noise)
    return err.max().item(), noise


def assert_close_scaled(got, want, what, tol=FP32_TOL, ulps=4):
    """north_star's 1e-5 is an absolute bar for O(1) quantities (frames, flows of a few pixels).  A gradient of
    magnitude 50 has an fp32 ulp of 3.8e-6: the bar for such values is 1e-5 + `ulps` ulp of the reference value."""
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    bound = tol + ulps * 2.0 ** -23 * want.abs()
    err = (got - want).abs()
    bad = err > bound
    assert not bad.any(), "%s: %d elements beyond %.0e + %d ulp (max abs err %.3e at |ref| %.3e)" % (
        what, int(bad.sum()), tol, ulps, err.max().item(), want.abs().flatten()[err.flatten().argmax()].item())
    return err.max().item()
