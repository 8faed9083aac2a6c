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
