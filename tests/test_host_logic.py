"""CPU suite: host-side logic -- synthetic inputs, padding rule, (pair, timestep) sharding incl. a
world_size-2 gloo run, coordinate-mode selection."""
import os
import subprocess
import sys

import pytest
import torch

import ssm_b200
from ssm_b200 import sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pad32_rule():
    # evaluate_interpolation_results.py:89-90: 720 -> 736, 1080 -> 1088, 2160 -> 2176, 352 stays
    assert [synthetic.pad32(n) for n in (720, 1080, 2160, 352, 1280, 1920)] == [736, 1088, 2176, 352, 1280, 1920]


def test_synthetic_inputs_are_seeded_and_in_range():
    a = synthetic.frames(2, 32, 48, seed=7)
    b = synthetic.frames(2, 32, 48, seed=7)
    assert a.shape == (2, 6, 32, 48) and torch.equal(a, b)
    assert a.min() >= -2.2 and a.max() <= 2.7           # ImageNet-normalised [0,1] frames
    f = synthetic.flows(2, 32, 48, 4, seed=8)
    assert f.shape == (2, 4, 32, 48) and f.abs().max() < 120
    assert synthetic.flows(1, 8, 8, 2, kind="zero").abs().max() == 0
    y = synthetic.unet_out5(2, 3, 32, 48)
    assert y.shape == (2, 3, 5, 32, 48)
    t = synthetic.timesteps(2, 7)
    assert torch.allclose(t[0], torch.arange(1, 8) / 8.0)
    assert torch.allclose(synthetic.timesteps(1, 31)[0, 0], torch.tensor(1 / 32.0))
    r = synthetic.random_timesteps(64)
    assert ((r * 8).round() == r * 8).all() and r.min() >= 0.125 and r.max() <= 0.875


@pytest.mark.parametrize("pairs,steps,world", [(16, 7, 1), (16, 7, 2), (16, 7, 8), (1, 31, 8), (3, 7, 8), (5, 7, 4)])
def test_shard_work_is_a_partition(pairs, steps, world):
    seen = set()
    sizes = []
    for r in range(world):
        w = sharding.shard_work(pairs, steps, r, world)
        sizes.append(sharding.frames_of(w))
        for p, t0, t1 in w:
            for t in range(t0, t1):
                assert (p, t) not in seen
                seen.add((p, t))
    assert len(seen) == pairs * steps
    assert max(sizes) - min(sizes) <= (steps if pairs >= world else 1)
    if pairs >= world:   # a pair's timesteps stay together
        for r in range(world):
            assert all(t0 == 0 and t1 == steps for _, t0, t1 in sharding.shard_work(pairs, steps, r, world))


def test_shard_work_c5_split():
    # one 4K pair, 31 timesteps over 8 ranks -> 4/4/4/4/4/4/4/3
    assert [sharding.frames_of(sharding.shard_work(1, 31, r, 8)) for r in range(8)] == [4] * 7 + [3]


def test_coord_mode_switch():
    assert ssm_b200.get_coord_mode() == 0
    ssm_b200.set_coord_mode("cuda")
    assert ssm_b200.get_coord_mode() == 1
    ssm_b200.set_coord_mode("cpu")
    assert ssm_b200.get_coord_mode() == 0
    with pytest.raises(KeyError):
        ssm_b200.set_coord_mode("cudnn")


_GLOO_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from ssm_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
work = sharding.shard_work(5, 7, rank, world)
mine = torch.zeros(5 * 7)
for p, t0, t1 in work:
    mine[p * 7 + t0: p * 7 + t1] = 1
dist.all_reduce(mine)              # harness-side check only: every work item owned exactly once
frames = torch.tensor([float(sharding.frames_of(work))])
dist.all_reduce(frames)
ok = bool((mine == 1).all()) and frames.item() == 35
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "gloo_shard.py"
    script.write_text(_GLOO_SCRIPT % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_frame_oracle_restatement_shapes_and_padding():
    """load_batch pads with byte 0 BEFORE normalising (visualize_interpolation.py:87,137): the border is
    (0 - mean) / std, not 0; the reader pads with zeros AFTER normalising (default_reader.py:266-271)."""
    import numpy as np
    from oracle import torch_oracle
    img = np.full((2, 100, 52, 3), 255, dtype=np.uint8)
    out = torch_oracle.load_batch_and_normalize(img)
    assert out.shape == (1, 2, 3, 128, 64)
    assert torch.allclose(out[0, 0, :, 0, 0], (0 - torch.tensor(torch_oracle.PIXEL_MEAN)) / torch.tensor(torch_oracle.PIXEL_STD))
    assert torch.allclose(out[0, 0, :, 14, 6], (1 - torch.tensor(torch_oracle.PIXEL_MEAN)) / torch.tensor(torch_oracle.PIXEL_STD))
    rd = torch_oracle.reader_normalize_and_pad(img, 8)
    assert rd.shape == (2, 3, 116, 52) and rd[:, :, :8].abs().max() == 0
    back = torch_oracle.crop_denormalize_u8(out[0], 14, 6, 100, 52)
    assert back.shape == (2, 100, 52, 3) and back.min() >= 254


def test_frames_host_helpers():
    import ssm_b200
    assert ssm_b200.frames.center_padding(720, 1280) == (736, 1280, 8, 0)
    assert ssm_b200.frames.center_padding(1080, 1920) == (1088, 1920, 4, 0)
    assert ssm_b200.frames.center_padding(100, 52) == (128, 64, 14, 6)
    lut = ssm_b200.normalisation_lut(device="cpu")
    assert lut.shape == (3, 256) and abs(lut[0, 255].item() - (1 - 0.485) / 0.229) < 1e-6
    with pytest.raises(RuntimeError):
        ssm_b200.frames_from_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))     # CPU tensor: refused


def test_sliding_window_matches_reference_fixture():
    """formats.sliding_window against the windows the reference's own Interpolator.sliding_window
    (scripts/visualize_interpolation.py:270-288) produced for 48 (n_images, n_frames, 240-fps) combinations
    (tests/golden/sliding_window.json, written by tests/golden/make_golden_large.py from the imported reference)."""
    import json
    import os
    from ssm_b200 import formats
    from util import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "sliding_window.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 48
    for c in cases:
        got = list(formats.sliding_window(c["n_images"], c["n_frames"], 8 if c["is_fps_240"] else 1))
        assert got == c["windows"], (c["n_images"], c["n_frames"], c["is_fps_240"])
    assert list(formats.sliding_window(4, 4)) == [[0, 0, 1, 2], [0, 1, 2, 3], [1, 2, 3, 3]]
    assert formats.output_name("d", 7) == "d/img_00007.png"


def test_flo_round_trip_and_header(tmp_path):
    """Middlebury .flo as scripts/utils/flo_utils.py:39-84 writes it: float 202021.25, int32 w, int32 h, data."""
    import numpy as np
    from ssm_b200 import formats
    flow = torch.randn(2, 5, 7)
    p = str(tmp_path / "a.flo")
    formats.write_flo(p, flow)
    raw = open(p, "rb").read()
    assert len(raw) == 12 + 5 * 7 * 2 * 4
    assert np.frombuffer(raw[:4], "<f4")[0] == np.float32(202021.25) and raw[:4] == b"PIEH"
    assert tuple(np.frombuffer(raw[4:12], "<i4")) == (7, 5)
    back = formats.read_flo(p)
    assert back.shape == (5, 7, 2) and np.array_equal(back, flow.permute(1, 2, 0).numpy())


def test_checkpoint_layout_round_trip(tmp_path):
    """reference layout: stage1_state_dict / stage2_state_dict in one file (main.py:231-237)."""
    from ssm_b200 import formats, unets

    class M:
        pass
    a, b = M(), M()
    for m, seed in ((a, 1), (b, 2)):
        torch.manual_seed(seed)
        m.stage1_model, m.stage2_model = unets.FlowUNet(6, 4, 1, True), unets.FlowUNet(16, 5, 2, True)
    p = str(tmp_path / "ckpt.pt")
    formats.save_checkpoint(a, p, iteration=12)
    assert set(torch.load(p)) >= {"stage1_state_dict", "stage2_state_dict"}
    assert formats.load_checkpoint(b, p) == 12
    for x, y in zip(a.stage2_model.parameters(), b.stage2_model.parameters()):
        assert torch.equal(x, y)


@pytest.mark.parametrize("bottleneck", ["CONV", "CLSTM", "CGRU"])
def test_state_dict_layout_matches_reference(bottleneck):
    """The in-tree U-Nets carry the reference's parameter names and shapes key for key, for every bottleneck the
    reference builds (tests/golden/state_dict_layout.json, dumped from the reference's own models by
    tests/golden/make_golden.py::dump_state_dict_layout): the author's checkpoints load unchanged."""
    import json
    import os
    from ssm_b200 import unets
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_dict_layout.json")) as f:
        want = json.load(f)[bottleneck]
    with torch.device("meta"):
        models = {"stage1": unets.FlowUNet(6, 4, 1, True, bottleneck=bottleneck),
                  "stage2": unets.FlowUNet(16, 5, 2, True, bottleneck=bottleneck)}
    for stage, m in models.items():
        got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert got == want[stage], stage


def test_get_model_reads_bottleneck_from_cfg():
    import configparser
    from ssm_b200 import recurrent, unets
    cfg = configparser.RawConfigParser()
    cfg.read_string("[STAGE1]\nBOTTLENECK=CGRU\n[STAGE2]\nBOTTLENECK=CONV\n")
    with torch.device("meta"):
        assert isinstance(unets.get_model(None, 6, 4, True, stage=1, cfg=cfg).conv6, recurrent.BiConvRecurrent)
        assert not isinstance(unets.get_model(None, 16, 5, True, stage=2, cfg=cfg).conv6, recurrent.BiConvRecurrent)
        cfg.set("STAGE2", "BOTTLENECK", "TRANSFORMER")
        with pytest.raises(AssertionError):
            unets.get_model(None, 16, 5, True, stage=2, cfg=cfg)


def test_numa_placement_helper(tmp_path, monkeypatch):
    """numa_local_to_gpu: binds to the GPU-local CPUs that the process may use, restores on exit, and is a
    no-op without topology information."""
    import os
    from ssm_b200 import sharding
    assert sharding._parse_cpulist("0-2,5,7-8\n") == {0, 1, 2, 5, 7, 8}
    before = os.sched_getaffinity(0)
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    some = sorted(before)[: max(1, len(before) // 2)]
    (d / "numa_node").write_text("0\n")
    (d / "local_cpulist").write_text(",".join(str(c) for c in some) + "\n")
    ctx = sharding.numa_local_to_gpu(0)
    monkeypatch.setattr(ctx, "_pci_dir", lambda: str(d))
    with ctx as placement:
        assert placement.info["numa_node"] == 0
        assert placement.info["cpus_bound"] == len(some)
        assert os.sched_getaffinity(0) == set(some)
    assert os.sched_getaffinity(0) == before
    # no topology (e.g. a VM without NUMA information): nothing happens
    (d / "numa_node").write_text("-1\n")
    ctx = sharding.numa_local_to_gpu(0)
    monkeypatch.setattr(ctx, "_pci_dir", lambda: str(d))
    with ctx as placement:
        assert placement.info == {"numa_node": None, "cpus_bound": 0, "mempolicy": False}
        assert os.sched_getaffinity(0) == before


class _RefStyleUNet(torch.nn.Module):
    """Built the way the reference builds its U-Nets (flow_computation.py:36-153): layers.conv Sequentials,
    layers.avg_pool, and the upsampleN lambdas stored as plain attributes."""

    def __init__(self):
        super().__init__()
        import torch.nn.functional as F
        self.conv1a, self.conv1b = ssm_b200.conv(6, 16, 7, padding=3), ssm_b200.conv(16, 16, 7, padding=3)
        self.pool2 = ssm_b200.avg_pool(kernel_size=2)
        self.conv2a = ssm_b200.conv(16, 32, 5, padding=2)
        self.conv6 = torch.nn.Sequential(ssm_b200.conv(32, 32), ssm_b200.conv(32, 32))
        self.upsample7 = lambda x: F.interpolate(x, size=(2 * x.shape[2], 2 * x.shape[3]), mode="bilinear")
        self.conv7a = ssm_b200.conv(48, 16)
        self.final_conv = torch.nn.Conv2d(16, 4, 3, padding=1)

    def forward(self, x):
        s1 = self.conv1b(self.conv1a(x))
        h = self.conv6(self.conv2a(self.pool2(s1)))
        return self.final_conv(self.conv7a(torch.cat([self.upsample7(h), s1], dim=1)))


def test_accelerate_unet_keeps_parameters_keys_and_cpu_results():
    from ssm_b200 import unet_glue
    torch.manual_seed(5)
    m = _RefStyleUNet()
    x = torch.randn(2, 6, 16, 24)
    want = m(x)
    keys = list(m.state_dict().keys())
    params = [id(p) for p in m.parameters()]
    assert ssm_b200.accelerate_unet(m) is m
    assert list(m.state_dict().keys()) == keys and [id(p) for p in m.parameters()] == params
    assert isinstance(m.conv1a, unet_glue.FusedConvLeaky) and isinstance(m.conv6[1], unet_glue.FusedConvLeaky)
    assert isinstance(m.pool2, unet_glue.FastAvgPool2) and m.upsample7 is unet_glue._fast_upsample
    assert isinstance(m.final_conv, torch.nn.Conv2d)
    got = m(x)                                  # CPU tensors: every fused module falls back to its stock children
    assert torch.allclose(got, want, atol=1e-6)
    got.sum().backward()
    assert all(p.grad is not None for p in m.parameters())
    ssm_b200.accelerate_unet(m)                 # idempotent
    assert list(m.state_dict().keys()) == keys


@pytest.mark.skipif(not os.path.isdir("/root/reference/scripts"), reason="needs the reference tree (build container only)")
def test_accelerate_unet_on_the_reference_classes():
    """The reference's own FlowComputationModel / FlowInterpolationModel: same keys, same CPU results after accelerate_unet."""
    import configparser
    sys.path.insert(0, "/root/reference/scripts")
    try:
        from models import unetflow
        cfg = configparser.RawConfigParser()
        cfg.read("/root/reference/configs/superslomo_original.ini")
        torch.manual_seed(6)
        s1 = unetflow.get_model(None, 6, 4, True, stage=1, cfg=cfg).eval()
        s2 = unetflow.get_model(None, 16, 5, True, stage=2, cfg=cfg).eval()
        x1, x2 = torch.randn(1, 1, 6, 64, 64), torch.randn(1, 1, 16, 64, 64)
        with torch.no_grad():
            o1 = s1(x1)
            o2 = s2(x2, [e for e, _ in o1])
        keys = (list(s1.state_dict()), list(s2.state_dict()))
        ssm_b200.accelerate_unet(s1); ssm_b200.accelerate_unet(s2)
        assert (list(s1.state_dict()), list(s2.state_dict())) == keys
        with torch.no_grad():
            p1 = s1(x1)
            p2 = s2(x2, [e for e, _ in p1])
        assert torch.allclose(p1[0][1], o1[0][1], atol=1e-5) and torch.allclose(p2[0], o2[0], atol=1e-5)
    finally:
        sys.path.remove("/root/reference/scripts")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]


def test_fullmodel_accepts_a_partial_config():
    """tools pass a config that only names the bottleneck: every other option keeps the reference's default."""
    import configparser
    cfg = configparser.RawConfigParser()
    cfg.read_string("[STAGE1]\nBOTTLENECK=CONV\n[STAGE2]\nBOTTLENECK=CONV\nCROSS_SKIP=TRUE\n")
    with torch.device("meta"):
        m = ssm_b200.FullModel(cfg=cfg, loss=ssm_b200.losses.SSMLosses(cfg, perceptual_features="zero"))
    assert m.loss.loss_weights == (60.0, 20.0, 10.0) and m.cross_skip


def test_perceptual_term_is_never_dropped_silently():
    """The reference always adds LAMBDA_P * VGG16 conv4_3 L2 (losses.py:12-41, 213).  With LAMBDA_P != 0 the loss module
    either has the pretrained extractor or refuses to be built; dropping the term takes an explicit opt-out, and an
    inference-only FullModel still constructs (the failure is deferred to the first training forward)."""
    import warnings
    from ssm_b200.losses import SSMLosses
    try:
        full = SSMLosses()                       # LAMBDA_P = 20: needs torchvision's pretrained VGG16
        assert full.perceptual_features is not None
    except RuntimeError as e:                    # offline
        assert "VGG16" in str(e) and "perceptual_features" in str(e)
        with torch.device("meta"):
            m = ssm_b200.FullModel(cfg=None)
        assert m.loss is None and "VGG16" in str(m._loss_error)
        with pytest.raises(RuntimeError, match="no loss module"):
            m(torch.zeros(1, 2, 3, 32, 32), torch.full((1, 1, 1, 1, 1), 0.5), target_images=torch.zeros(1, 1, 3, 32, 32),
              inference_mode=False)
    assert SSMLosses(perceptual_features="zero").perceptual_features is None
    assert SSMLosses(lambda_p=0.0).perceptual_features is None            # nothing to drop
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        SSMLosses(perceptual_features=None)
    assert any("DROPPED" in str(x.message) for x in w)
    rnd = SSMLosses(perceptual_features="random")
    assert len(rnd.perceptual_features) == 23 and not any(p.requires_grad for p in rnd.perceptual_features.parameters())


@pytest.mark.skipif(not os.path.isdir("/root/reference/scripts"), reason="needs the reference tree (build container only)")
def test_patch_reference_rebinds_the_reference_modules():
    """ssm_b200.patch_reference on the REAL reference modules: the six rebindings INTEGRATION.md section 1 promises
    (three methods of FlowInterpolationModel, `warp` in flow_interpolation / layers / losses), idempotence, and that CPU
    tensors still reach the reference's own functions bit for bit (the patched names route by device)."""
    import importlib
    import sys
    import ssm_b200
    from ssm_b200 import synthetic
    sys.path.insert(0, "/root/reference/scripts")
    saved = {k: sys.modules.get(k) for k in ("models", "models.flow_interpolation", "models.layers", "models.losses")}
    try:
        fi = importlib.import_module("models.flow_interpolation")
        layers = importlib.import_module("models.layers")
        try:
            losses = importlib.import_module("models.losses")
        except Exception:                      # torchvision / VGG import problems are not what this test is about
            losses = None
        cls = fi.FlowInterpolationModel
        originals = {n: getattr(cls, n) for n in ("compute_inputs", "extract_outputs", "compute_output_image")}
        ref_warp = layers.warp
        assert fi.warp is ref_warp
        ssm_b200.patch_reference(fi, layers, losses)
        rebound = 0
        for n, orig in originals.items():
            assert getattr(cls, n) is not orig and getattr(getattr(cls, n), "_ssm_b200", False), n
            assert getattr(cls, "_ssm_ref_" + n) is orig
            rebound += 1
        for mod in (fi, layers, losses):
            if mod is None:
                continue
            assert mod.warp is not ref_warp and mod.warp._ssm_b200 and mod.warp._ssm_ref is ref_warp
            rebound += 1
        assert rebound == (6 if losses is not None else 5)
        for n in ("compute_inputs_batched", "compute_output_image_batched", "compute_output_image_from_flow"):
            assert hasattr(cls, n)
        before = {n: getattr(cls, n) for n in originals}
        ssm_b200.patch_reference(fi, layers, losses)                  # idempotent: no wrapper around a wrapper
        assert all(getattr(cls, n) is before[n] for n in originals) and layers.warp._ssm_ref is ref_warp

        # CPU tensors: the reference's own code path, unchanged
        class Bare:
            verbose = False
        bare = Bare()
        img6 = synthetic.frames(1, 16, 24, seed=11)
        flow4 = synthetic.flows(1, 16, 24, 4, flow_px=3.0, seed=12)
        out5 = synthetic.unet_out5(1, 1, 16, 24, seed=13)[:, 0].contiguous()
        t = torch.tensor([0.25]).view(1, 1, 1, 1)
        bare.extract_outputs = lambda y: cls.extract_outputs(bare, y)
        want16 = originals["compute_inputs"](bare, img6, flow4, t)
        assert torch.equal(cls.compute_inputs(bare, img6, flow4, t), want16)
        want3 = originals["compute_output_image"](bare, img6, want16, out5, t)
        assert torch.equal(cls.compute_output_image(bare, img6, want16, out5, t), want3)
        assert torch.equal(layers.warp(img6[:, :3], flow4[:, :2]), ref_warp(img6[:, :3], flow4[:, :2]))
    finally:
        for k in ("models.flow_interpolation", "models.layers", "models.losses", "models"):
            if saved[k] is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = saved[k]
        sys.path.remove("/root/reference/scripts")


def test_device_t_check_modes():
    import ssm_b200
    assert ssm_b200.set_device_t_check("flag") in ("off", "flag", "assert")
    assert ssm_b200.set_device_t_check("off") == "flag"
    with pytest.raises(ValueError):
        ssm_b200.set_device_t_check("sometimes")
    assert ssm_b200.t_violations() == 0
