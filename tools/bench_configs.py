"""The other BASELINE.json configurations, run by bench.py next to the headline (configs[1]) so that the driver's
1/2/4/8-GPU runs record them:

  C3  configs[2]  SuperSloMo training step, 352x352 crops, GLOBAL batch 64, data-parallel: one process per GPU,
                  DistributedDataParallel over NCCL (the reference is single-process nn.DataParallel,
                  scripts/main.py:74-76, 185-186).  Strong scaling: 64 / N samples per GPU.  Reports ms/step, samples/s,
                  the all-reduced bytes, the EXPOSED communication time (against the same per-GPU step on an unwrapped
                  copy of the model) and the share of the step spent in this repo's path kernels.
  C4  configs[3]  superslomo_recurrent.ini (SSMR: N_FRAMES = 4 -> 3 windows, bidirectional ConvLSTM bottleneck,
                  configs/superslomo_recurrent.ini:82, 97, 105) on 1088x1920 sequences, 7 intermediate times of the middle
                  window, one sequence per GPU (weak scaling: the ConvLSTM couples the windows of a sample, so samples
                  are the unit that shards).
  C5  configs[4]  one 2176x3840 (4K padded to /32) pair x 31 intermediate times (t = k/32,
                  scripts/evaluate_interpolation_results.py:52, 204-211) split 4/4/4/4/4/4/4/3 over the ranks
                  (sharding.shard_work), INCLUDING the NCCL broadcast of the pair (two 8-bit images) and its stage-1 flows
                  from rank 0; every rank then runs the 8-bit-frame kernels on its timesteps.

The two flow U-Nets are stock torch/cuDNN modules with random-init weights (out of scope); the
timed work of C5 is the path alone (stage-2 output = seeded surrogate, as in the headline).  Every number is
device-timed with CUDA events and reduced with MAX over ranks.
"""
import configparser
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PATH_ENTRY_POINTS = ["ssm_pack_frames", "ssm_flow_pack_fwd", "ssm_flow_pack_bwd", "ssm_fuse_loss_fwd", "ssm_fuse_loss_bwd",
                     "ssm_fuse_flow_fwd", "ssm_fuse_flow_bwd", "ssm_fuse_fwd", "ssm_fuse_bwd", "ssm_warp_fwd", "ssm_warp_bwd",
                     "ssm_flow_pack_fwd_nhwc", "ssm_fuse_flow_fwd_mixed", "ssm_quads_from_u8", "ssm_flow_pack_fwd_q8",
                     "ssm_flow_pack_fwd_q8_nhwc", "ssm_flow_pack_fwd_q8_lut", "ssm_fuse_flow_fwd_q8", "ssm_fuse_flow_fwd_q8_u8", "ssm_frames_from_u8",
                     "ssm_frames_to_u8"]


class PathTimer:
    """Events around every C-ABI call of the path while active: {entry point: ms}."""

    def __init__(self):
        import ssm_b200
        self.abi = ssm_b200._abi
        self.spans = []

    def __enter__(self):
        lib = self.abi.lib()
        spans = self.spans

        class Shim:          # ctypes function objects cannot be replaced on the CDLL: shim the attribute lookup
            def __getattr__(self, n):
                fn = getattr(lib, n)
                if n not in PATH_ENTRY_POINTS:
                    return fn

                def call(*args):
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    rc = fn(*args)
                    e.record()
                    spans.append((n, s, e))
                    return rc
                return call
        self._lib = lib
        self.abi._lib = Shim()
        return self

    def __exit__(self, *exc):
        self.abi._lib = self._lib
        return False

    def result(self):
        torch.cuda.synchronize()
        out = {}
        for n, s, e in self.spans:
            out[n] = out.get(n, 0.0) + s.elapsed_time(e)
        return out


def _max_over_ranks(value, dev, world):
    if world > 1:
        t = torch.tensor([float(value)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    return float(value)


def _time_steps(step, n, dev, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return _max_over_ranks(e0.elapsed_time(e1) / n, dev, world)


# ---------------------------------------------------------------------------------------------
def c3_train_step(world, rank, dev, steps=5, warmup=3, global_batch=64, size=352):
    import ssm_b200
    from ssm_b200 import synthetic
    from ssm_b200.superslomo_r import FullModel
    if global_batch % world != 0:
        return {"skipped": "global batch %d does not divide over %d ranks" % (global_batch, world)}
    B = global_batch // world
    torch.manual_seed(42)                        # same initial weights on every rank (configs/*.ini [SEED])
    cfg = configparser.RawConfigParser()
    cfg.read_string("[STAGE1]\nBOTTLENECK=CONV\n[STAGE2]\nBOTTLENECK=CONV\nCROSS_SKIP=TRUE\n")
    # LAMBDA_R/P/W of configs/superslomo_original.ini; the VGG16 of the perceptual term is random-init (no network):
    # same compute as the reference's pretrained one
    loss = ssm_b200.losses.SSMLosses(lambda_r=60.0, lambda_p=20.0, lambda_w=10.0, perceptual_features="random")
    model = FullModel(cfg=cfg, loss=loss).to(dev)
    model.stage1_model.set_channels_last()
    model.stage2_model.set_channels_last()
    model.loss.perceptual_features.to(memory_format=torch.channels_last)
    bucket_mb = 25
    # a second copy of the model WITHOUT the DDP wrapper: the same step with no communication at all, the yardstick for
    # the exposed all-reduce time (DDP.no_sync() is not one: it changes how gradients are accumulated -- at 8 GPUs the
    # no_sync step measured SLOWER than the synchronised one)
    local = copy.deepcopy(model) if world > 1 else None
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], bucket_cap_mb=bucket_mb,
                                                    gradient_as_bucket_view=True) if world > 1 else model
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    n_params = sum(p.numel() for p in params)
    frames = synthetic.frames(B, size, size, n_frames=2, seed=100 + rank, device=dev).view(B, 2, 3, size, size)
    targets = synthetic.frames(B, size, size, n_frames=1, seed=200 + rank, device=dev).view(B, 1, 3, size, size)
    t = synthetic.random_timesteps(B, 1, seed=300 + rank).to(dev).view(B, 1, 1, 1, 1)

    def make_step(module, optimizer):
        def step():
            optimizer.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                _, losses = module(frames, t, target_images=targets, inference_mode=False)
            losses[:, 0].float().mean().backward()
            optimizer.step()
        return step

    step = make_step(net, opt)
    for _ in range(warmup):
        step()
    ms = _time_steps(step, steps, dev, world)
    ms_local = ms
    if world > 1:
        step_local = make_step(local, torch.optim.Adam([p for p in local.parameters() if p.requires_grad], lr=1e-4))
        for _ in range(warmup):
            step_local()
        ms_local = _time_steps(step_local, steps, dev, world)
        del step_local, local
    with PathTimer() as pt:
        step()
    path = pt.result()
    path_total = sum(path.values())
    res = {
        "what": "SuperSloMo training step, %dx%d crops, global batch %d, fwd + bwd + Adam, bf16-autocast channels-last "
                "U-Nets (stock torch/cuDNN, random init), fp32 path kernels, perceptual term with a random-init VGG16" % (size, size, global_batch),
        "parallelism": "DistributedDataParallel over NCCL, one process per GPU (reference: nn.DataParallel, scripts/main.py:74-76)"
                       if world > 1 else "single GPU",
        "scaling": "strong", "n_gpus": world, "global_batch": global_batch, "per_gpu_batch": B,
        "ms_per_step": ms, "samples_per_s": global_batch / (ms * 1e-3),
        "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0, "ddp_bucket_mb": bucket_mb if world > 1 else None,
        "ms_per_step_without_allreduce": ms_local, "exposed_comm_ms": max(ms - ms_local, 0.0) if world > 1 else 0.0,
        "exposed_comm_how": "the same per-GPU step on an unwrapped copy of the model (no DDP hooks, no collective), max over ranks",
        "trainable_parameters": n_params,
        "path_kernels_ms": path, "path_ms_total": path_total, "path_share_of_step": path_total / ms,
    }
    del net, model, opt, frames, targets
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------
def c4_ssmr_windows(world, rank, dev, steps=2, warmup=1, H=1088, W=1920, n_frames=4, n_t=7):
    import ssm_b200  # noqa: F401
    from ssm_b200 import synthetic
    from ssm_b200.superslomo_r import FullModel
    torch.manual_seed(42)
    cfg = configparser.RawConfigParser()
    cfg.read_string("[STAGE1]\nBOTTLENECK=CLSTM\n[STAGE2]\nBOTTLENECK=CLSTM\nCROSS_SKIP=TRUE\n")
    model = FullModel(cfg=cfg, loss=ssm_b200.losses.SSMLosses(perceptual_features="zero")).to(dev).eval()
    model.stage1_model.set_channels_last()
    model.stage2_model.set_channels_last()
    B = 1
    H_in = 1080 if H == 1088 else H                 # 8-bit 1080 x 1920 images, centred in 1088 x 1920 by the library
    x = synthetic.frames(B * n_frames, H_in, W, n_frames=1, seed=400 + rank, smooth=True, device=dev)
    x = (x - x.amin()) / (x.amax() - x.amin())
    clip = (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).view(B, n_frames, H_in, W, 3).contiguous()
    del x
    t_values = torch.tensor([(k + 1) / (n_t + 1) for k in range(n_t)], dtype=torch.float32, device=dev)

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model.interpolate_u8(clip, t_values, order="rgb", unet_chunk=1)

    for _ in range(warmup):
        step()
    ms = _time_steps(step, steps, dev, world)
    with PathTimer() as pt:
        step()
    path = pt.result()
    path_total = sum(path.values())
    res = {
        "what": "superslomo_recurrent.ini (SSMR): %d-frame windows (%d windows, bidirectional ConvLSTM bottleneck), "
                "%dx%d (8-bit images in, 8-bit interpolated images out: FullModel.interpolate_u8), %d intermediate times of the "
                "middle window per sequence, inference, bf16-autocast channels-last U-Nets (stock torch/cuDNN, random init)"
                % (n_frames, n_frames - 1, H, W, n_t),
        "parallelism": "one sequence per GPU, no collective (the ConvLSTM couples the windows of a sequence)",
        "scaling": "weak", "n_gpus": world, "sequences_per_gpu": B,
        "ms_per_step": ms, "frames_per_s": B * n_t * world / (ms * 1e-3),
        "path_kernels_ms": path, "path_ms_total": path_total, "path_share_of_step": path_total / ms,
    }
    del model, clip
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------
def c5_4k_sharded(world, rank, dev, steps=5, warmup=3, H_in=2160, W_in=3840, n_t=31, peak_gbs=None):
    import ssm_b200
    from ssm_b200 import q8, sharding, synthetic
    H, W = (H_in + 31) // 32 * 32, (W_in + 31) // 32 * 32          # 2176 x 3840 (evaluate_interpolation_results.py:89-90)
    work = sharding.shard_work(1, n_t, rank, world)
    assert len(work) == 1
    _, t0, t1 = work[0]
    n = t1 - t0
    # the pair lives on rank 0 as the two 8-bit images a video decoder delivers, next to its stage-1 flows
    if rank == 0:
        x = synthetic.frames(2, H_in, W_in, n_frames=1, seed=500, smooth=True, device=dev)
        x = (x - x.amin()) / (x.amax() - x.amin())
        images = (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
        del x
        flow4 = synthetic.flows(1, H, W, 4, flow_px=20.0, seed=501, device=dev)
    else:
        images = torch.empty((2, H_in, W_in, 3), dtype=torch.uint8, device=dev)
        flow4 = torch.empty((1, 4, H, W), device=dev)
    lut = ssm_b200.normalisation_lut(device=dev)
    pads = lut[:, 0].tolist()
    t_all = synthetic.timesteps(1, n_t, device=dev)
    t = t_all[:, t0:t1].contiguous()
    out5 = synthetic.unet_out5(1, n, H, W, seed=502 + rank, device=dev)
    in16 = torch.empty((1, n, 16, H, W), device=dev)
    frames = torch.empty((1, n, 3, H, W), device=dev)
    ev = []

    def step():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        if world > 1:                         # the pair and its stage-1 flows live on rank 0
            dist.broadcast(images, src=0)
            dist.broadcast(flow4, src=0)
        e[1].record()
        with torch.no_grad():
            planar, quads, norm, _ = q8.prepare(images, order="rgb", lut=lut, pad_values=pads)
            q8.flow_pack(None, quads, flow4, t, norm, n_timesteps=n, out=in16, lut=lut)     # pass-through channels from the tables
            q8.fuse_from_flow(quads, flow4, out5, t, norm, out=frames)
        e[2].record()
        ev.append(e)

    for _ in range(warmup):
        step()
    del ev[:]
    ms = _time_steps(step, steps, dev, world)
    bcast = _max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in ev) / len(ev), dev, world)
    kern = _max_over_ranks(sum(e[1].elapsed_time(e[2]) for e in ev) / len(ev), dev, world)
    npx = H * W
    nbytes = ((10 + 16 * n) + (10 + 8 * n)) * 4 * npx          # this rank's algorithmic bytes
    res = {
        "what": "one %dx%d pair (two 8-bit %dx%d images) x %d intermediate times (t = k/32), (pair, timestep) work split over "
                "the ranks; fp32 tensors; step = NCCL broadcast of the two uint8 images and the stage-1 flows (4 fp32 planes) from "
                "rank 0 + frame normalisation and entry tables + compute_inputs + compute_output_image for this rank's "
                "timesteps" % (H, W, H_in, W_in, n_t),
        "parallelism": "timesteps of the single pair split %s, one NCCL broadcast per step" % "/".join(
            str(sharding.frames_of(sharding.shard_work(1, n_t, r, world))) for r in range(world)),
        "scaling": "strong", "n_gpus": world, "timesteps_this_rank": n,
        "ms_per_step": ms, "frames_per_s": n_t / (ms * 1e-3),
        "broadcast_ms": bcast, "broadcast_bytes": (images.numel() + 4 * 4 * npx) if world > 1 else 0, "path_kernels_ms": kern,
        "rank0_path_algorithmic_gbs": nbytes / (kern * 1e-3) / 1e9,
    }
    if peak_gbs:
        res["rank0_path_frac_of_peak"] = res["rank0_path_algorithmic_gbs"] / peak_gbs
    del images, flow4, out5, in16, frames
    torch.cuda.empty_cache()
    return res


def run_all(world, rank, dev, peak_gbs=None):
    out = {}
    for name, fn in (("C3_train_step_352_b64", lambda: c3_train_step(world, rank, dev)),
                     ("C4_ssmr_windows_1080p", lambda: c4_ssmr_windows(world, rank, dev)),
                     ("C5_4k_31_timesteps", lambda: c5_4k_sharded(world, rank, dev, peak_gbs=peak_gbs))):
        try:
            out[name] = fn()
        except Exception as e:        # a failing side configuration must not take the headline line down: say what failed
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.empty_cache()
            if world > 1:
                raise                 # ranks must stay in lock step: a one-sided failure would hang the others
    return out
