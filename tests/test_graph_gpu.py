"""CUDA-graph capture of the path: the C ABI never allocates, frees or
synchronises and takes the caller's stream, so a whole N-timestep step -- frame pre-pass, compute_inputs,
compute_output_image, frame post-pass -- is capturable and replays bit-identically with new inputs."""
import numpy as np
import pytest
import torch

import ssm_b200
from ssm_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B, N, H, W, seed):
    img6 = synthetic.frames(B, H, W, seed=seed, device=DEV)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=6.0, seed=seed + 1, device=DEV)
    out5 = synthetic.unet_out5(B, N, H, W, seed=seed + 2, device=DEV)
    return img6, flow4, out5


def test_step_is_capturable_and_replays_with_new_inputs():
    B, N, H, W = 2, 3, 64, 96
    t = synthetic.timesteps(B, N, device=DEV)
    img6, flow4, out5 = (x.clone() for x in _inputs(B, N, H, W, 1))          # static input buffers

    def step():
        rgbx = ssm_b200.pack_frames(img6)
        in16 = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=N, packed=rgbx)
        return in16, ssm_b200.fuse_from_flow(img6, flow4, out5, t, packed=rgbx)

    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_in16, g_frames = step()
        for seed in (1, 50):
            a, b, c = _inputs(B, N, H, W, seed)
            img6.copy_(a); flow4.copy_(b); out5.copy_(c)
            graph.replay()
            torch.cuda.synchronize()
            want_in16, want_frames = step()
            assert torch.equal(g_in16, want_in16) and torch.equal(g_frames, want_frames)


def test_frame_kernels_are_capturable():
    rng = np.random.RandomState(0)
    u8 = torch.from_numpy(rng.randint(0, 256, size=(2, 45, 70, 3)).astype(np.uint8)).to(DEV)
    lut = ssm_b200.normalisation_lut(device=DEV)
    pad_values = lut[:, 0].tolist()

    def step():
        planar, rgbx, (top, left) = ssm_b200.frames_from_u8(u8, lut=lut, want_rgbx=True, pad_values=pad_values)
        return planar, ssm_b200.frames_to_u8(planar, top=top, left=left, h_out=45, w_out=70, order="bgr", saturate=True)

    want_planar, want_back = step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        planar, back = step()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(planar, want_planar) and torch.equal(back, want_back)
