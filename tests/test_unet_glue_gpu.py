"""GPU suite: the element-wise kernels between the U-Nets' cuDNN convolutions (ssm_upsample2x_nhwc,
ssm_bias_leaky_nhwc, ssm_avgpool2_nhwc) against the ATen ops they replace, and the U-Net inference fast path
against the stock torch modules.  fp32: same expression, same order -> at most one ulp apart (FMA contraction is
the compiler's choice on both sides); bf16: at most one bf16 ulp (the single final rounding)."""
import pytest
import torch
import torch.nn.functional as F

import ssm_b200
from ssm_b200 import unet_glue
from util import seeded_unets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nhwc(M, C, H, W, dtype, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((M, C, H, W), generator=g, device=DEV) * 3.0
    return x.to(dtype).contiguous(memory_format=torch.channels_last)


def _ulp_bound(ref, dtype, mag):
    """mag = the same op applied to |x|: an upper bound of every intermediate sum.  fp32: the two sides may
    contract different products into FMAs (2 ulp of the operands' magnitude); bf16: that, plus one ulp of the
    result when the difference flips the single final rounding."""
    bound = 2.0 ** -22 * mag.float()
    if dtype == torch.bfloat16:
        bound = bound + 2.0 ** -8 * ref.float().abs()
    return bound


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C,H,W", [(2, 8, 1, 1), (1, 16, 1, 7), (1, 24, 5, 1), (3, 64, 17, 34), (1, 128, 68, 120)])
def test_upsample2x_matches_aten(dtype, M, C, H, W):
    x = _nhwc(M, C, H, W, dtype, 1)
    ref = F.interpolate(x, size=(2 * H, 2 * W), mode="bilinear", align_corners=False)
    got = unet_glue.upsample2x_cat([x])
    mag = F.interpolate(x.float().abs(), size=(2 * H, 2 * W), mode="bilinear", align_corners=False)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert ((got.float() - ref.float()).abs() <= _ulp_bound(ref, dtype, mag)).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upsample2x_of_a_concatenation(dtype):
    a, b, c = _nhwc(2, 32, 9, 13, dtype, 2), _nhwc(2, 8, 9, 13, dtype, 3), _nhwc(2, 64, 9, 13, dtype, 4)
    ref = F.interpolate(torch.cat([a, b, c], dim=1), size=(18, 26), mode="bilinear", align_corners=False)
    got = unet_glue.upsample2x_cat([a, b, c])
    mag = F.interpolate(torch.cat([a, b, c], dim=1).float().abs(), size=(18, 26), mode="bilinear", align_corners=False)
    assert ((got.float() - ref.float()).abs() <= _ulp_bound(ref, dtype, mag)).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upsample2x_propagates_non_finite_like_aten(dtype):
    x = _nhwc(1, 8, 4, 4, dtype, 5)
    x[0, 0, 1, 1] = float("inf")
    x[0, 1, 0, 0] = float("nan")
    ref = F.interpolate(x, size=(8, 8), mode="bilinear", align_corners=False).float()
    got = unet_glue.upsample2x_cat([x]).float()
    assert torch.equal(torch.isnan(got), torch.isnan(ref)) and torch.equal(torch.isinf(got), torch.isinf(ref))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C,H,W", [(1, 8, 2, 2), (2, 32, 12, 20), (1, 512, 34, 60)])
def test_avgpool2_matches_aten(dtype, M, C, H, W):
    x = _nhwc(M, C, H, W, dtype, 6)
    ref = F.avg_pool2d(x, 2)
    got = unet_glue.avgpool2(x)
    assert got.shape == ref.shape
    assert ((got.float() - ref.float()).abs() <= _ulp_bound(ref, dtype, F.avg_pool2d(x.float().abs(), 2))).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bias_leaky_matches_the_two_aten_ops(dtype):
    y = _nhwc(2, 32, 11, 19, dtype, 7)
    bias = torch.randn(32, device=DEV)
    b = bias.to(dtype)
    ref = F.leaky_relu(y + b.view(1, -1, 1, 1), 0.1)
    got = unet_glue.bias_leaky_(y.clone(memory_format=torch.preserve_format), b.float(), 0.1)
    assert ((got.float() - ref.float()).abs() <= _ulp_bound(ref, dtype, y.float().abs() + b.float().abs().view(1, -1, 1, 1))).all()


def test_glue_rejects_what_it_cannot_do():
    x = _nhwc(1, 8, 4, 4, torch.float32, 8)
    assert unet_glue.usable(x)
    assert not unet_glue.usable(x.contiguous())                              # planar
    assert not unet_glue.usable(_nhwc(1, 12, 4, 4, torch.float32, 9))        # C % 8
    assert not unet_glue.usable(x.half())
    assert not unet_glue.usable(x.cpu())
    L = ssm_b200._abi.lib()
    import ctypes
    rc = L.ssm_upsample2x_nhwc(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(x.data_ptr()), 1, 4, 4, 12, 12, 0, None)
    assert rc < 0 and b"multiple of 8" in L.ssm_last_error()


@pytest.mark.parametrize("amp", [False, True])
@pytest.mark.parametrize("bottleneck", ["CONV", "CLSTM"])
def test_unet_fast_path_matches_stock_modules(amp, bottleneck):
    """Both U-Nets, channels-last, inference: element-wise steps on this repo's kernels vs stock torch ops."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        s1, s2 = seeded_unets(31, DEV, bottleneck=bottleneck)
        s1.eval().set_channels_last(); s2.eval().set_channels_last()
        g = torch.Generator(device=DEV).manual_seed(11)
        pairs = torch.randn((2, 3, 6, 64, 96), generator=g, device=DEV)
        in16 = torch.randn((2, 3, 16, 64, 96), generator=g, device=DEV)
        outs = {}
        for fast in (False, True):
            s1.fast_glue = s2.fast_glue = fast
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                o1 = s1(pairs)
                o2 = s2(in16, [e for e, _ in o1])
            outs[fast] = (torch.stack([f for _, f in o1]).float(), torch.stack(o2).float())
        # fp32: ulp-level differences in the element-wise steps, amplified by ~20 convolutions; bf16: every
        # activation is re-rounded to 8 bits, so a one-ulp difference anywhere moves the output by ~1 %
        tol = 3e-2 if amp else 1e-4
        for a, b in zip(outs[False], outs[True]):
            scale = a.abs().max().item()
            assert (a - b).abs().max().item() <= tol * max(scale, 1.0), (amp, bottleneck)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_backward_kernels_match_autograd_of_the_aten_ops(dtype):
    """upsample(cat), avg-pool and bias + LeakyReLU: gradients from the gather kernels vs autograd through the ATen ops."""
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    a, b = _nhwc(2, 16, 7, 9, dtype, 21).requires_grad_(), _nhwc(2, 8, 7, 9, dtype, 22).requires_grad_()
    g = _nhwc(2, 24, 14, 18, dtype, 23)
    ref = F.interpolate(torch.cat([a, b], dim=1), size=(14, 18), mode="bilinear", align_corners=False)
    ra, rb = torch.autograd.grad(ref, (a, b), g)
    ga, gb = torch.autograd.grad(unet_glue.upsample2x_cat([a, b]), (a, b), g)
    for got, want in ((ga, ra), (gb, rb)):
        assert (got.float() - want.float()).abs().max().item() <= tol * max(1.0, want.float().abs().max().item())
    for shape in ((1, 8, 1, 1), (1, 8, 1, 5), (2, 8, 6, 1)):          # degenerate sizes: every border rule at once
        x = _nhwc(*shape, dtype, 24).requires_grad_()
        gg = _nhwc(shape[0], shape[1], 2 * shape[2], 2 * shape[3], dtype, 25)
        (want,) = torch.autograd.grad(F.interpolate(x, size=(2 * shape[2], 2 * shape[3]), mode="bilinear", align_corners=False), x, gg)
        (got,) = torch.autograd.grad(unet_glue.upsample2x_cat([x]), x, gg)
        assert (got.float() - want.float()).abs().max().item() <= tol * max(1.0, want.float().abs().max().item()), shape
    x = _nhwc(2, 32, 10, 12, dtype, 26).requires_grad_()
    gp = _nhwc(2, 32, 5, 6, dtype, 27)
    (want,) = torch.autograd.grad(F.avg_pool2d(x, 2), x, gp)
    (got,) = torch.autograd.grad(unet_glue.avgpool2(x), x, gp)
    assert (got.float() - want.float()).abs().max().item() <= tol * max(1.0, want.float().abs().max().item())
    y0 = _nhwc(2, 32, 9, 11, dtype, 28)
    bias = torch.randn(32, device=DEV, requires_grad=True)
    gy = _nhwc(2, 32, 9, 11, dtype, 29)
    ya = y0.clone(memory_format=torch.preserve_format).requires_grad_()
    ref = F.leaky_relu(ya + bias.to(dtype).view(1, -1, 1, 1), 0.1)
    wy, wb = torch.autograd.grad(ref, (ya, bias), gy)
    yb = y0.clone(memory_format=torch.preserve_format).requires_grad_()
    out = unet_glue.bias_leaky_(yb * 1.0, bias, 0.1)                    # * 1.0: a non-leaf, as a convolution output is
    qy, qb = torch.autograd.grad(out, (yb, bias), gy)
    assert (qy.float() - wy.float()).abs().max().item() <= tol * max(1.0, wy.float().abs().max().item())
    assert (qb.float() - wb.float()).abs().max().item() <= (1e-3 if dtype == torch.float32 else 5e-2) * max(1.0, wb.abs().max().item())


@pytest.mark.parametrize("bottleneck", ["CONV", "CLSTM"])
def test_unet_training_fast_path_matches_stock_modules(bottleneck):
    """fp32, channels-last, autograd recording: outputs and every parameter gradient, fast vs stock element-wise ops."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        s1, s2 = seeded_unets(33, DEV, bottleneck=bottleneck)
        s1.set_channels_last(); s2.set_channels_last()
        g = torch.Generator(device=DEV).manual_seed(12)
        pairs = torch.randn((2, 2, 6, 64, 64), generator=g, device=DEV)
        in16 = torch.randn((2, 2, 16, 64, 64), generator=g, device=DEV)
        res = {}
        for fast in (False, True):
            s1.fast_glue = s2.fast_glue = fast
            s1.zero_grad(); s2.zero_grad()
            o1 = s1(pairs)
            o2 = s2(in16, [e for e, _ in o1])
            loss = sum(f.square().mean() for _, f in o1) + sum(o.square().mean() for o in o2)
            loss.backward()
            res[fast] = (loss.item(), {n: p.grad.clone() for m in (s1, s2) for n, p in m.named_parameters()})
        assert abs(res[True][0] - res[False][0]) <= 1e-4 * max(1.0, abs(res[False][0]))
        for n, want in res[False][1].items():
            got = res[True][1][n]
            assert (got - want).abs().max().item() <= 2e-3 * max(want.abs().max().item(), 1e-6), n
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("amp", [False, True])
def test_accelerate_unet_on_a_reference_style_module(amp):
    """accelerate_unet on a U-Net built the way the reference builds its own (layers.conv Sequentials, avg_pool,
    upsampleN lambdas): same results as the untouched module, and the element-wise kernels are what runs."""
    import copy
    from test_host_logic import _RefStyleUNet
    torch.manual_seed(7)
    stock = _RefStyleUNet().to(DEV).eval()
    fast = ssm_b200.accelerate_unet(copy.deepcopy(stock))
    x = torch.randn(2, 6, 32, 48, device=DEV)
    calls = []
    lib = ssm_b200._abi.lib()

    class _Spy:
        def __getattr__(self, n):
            if n in ("ssm_upsample2x_nhwc", "ssm_bias_leaky_nhwc", "ssm_avgpool2_nhwc"):
                calls.append(n)
            return getattr(lib, n)

    ssm_b200._abi._lib = _Spy()
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            want, got = stock(x).float(), fast(x).float()
    finally:
        ssm_b200._abi._lib = lib
    assert calls.count("ssm_bias_leaky_nhwc") == 6 and calls.count("ssm_avgpool2_nhwc") == 1 and calls.count("ssm_upsample2x_nhwc") == 1
    tol = 3e-2 if amp else 1e-4
    assert (got - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    # training through the accelerated module: gradients agree with the stock one
    stock.train(); fast.train()
    ls, lf = stock(x).square().mean(), fast(x).square().mean()
    ls.backward(); lf.backward()
    for (n, a), (_, b) in zip(stock.named_parameters(), fast.named_parameters()):
        assert (a.grad - b.grad).abs().max().item() <= 2e-3 * max(a.grad.abs().max().item(), 1e-6), n
