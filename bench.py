#!/usr/bin/env python
"""bench.py -- throughput of the intermediate-frame synthesis path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 16 synthetic 1080p frame pairs per GPU, 7 intermediate timesteps.  The frames are
8-bit images (what a decoded video frame is and what the reference reads, scripts/visualize_interpolation.py:61-72),
1080x1920 centred in 1088x1920 (padded to /32 with byte 0, :76-87), normalised with the reference's expression.
One step = one pass of the hot path over the batch:
    ssm_quads_from_u8        staging: the 2x2 byte entry tables the gathers read (one launch; charged to the step's time,
                             not to its bytes)
    ssm_flow_pack_fwd_q8     compute_inputs for all 7 timesteps (one launch)
    ssm_fuse_flow_fwd_q8     extract_outputs + compute_output_image for all 7 timesteps (one launch)
= 112 interpolated frames per GPU.  The two flow U-Nets are out of scope (they stay on PyTorch/cuDNN); the stage-1
flows and the stage-2 output are seeded surrogates.

  value        frames/s, inputs resident in HBM, device-timed with CUDA events, max over ranks
  e2e          same metric through ssm_synthesize_host_u8: pinned HOST buffers (uint8 frames, fp32 flows, bf16 U-Net output)
               in, uint8 interpolated images out, host<->device copies inside the timed region
  roofline     dominant kernel (compute_inputs: 65 % of the path's bytes), algorithmic bytes / event-timed duration vs the
               measured HBM peak; roofline_kernels lists every kernel of the path, both frame formats
  fp32_frames_path  the same step for callers that hold normalised fp32 frames (the reference's tensor interface and the
               round-1 headline): ssm_pack_frames + ssm_flow_pack_fwd + ssm_fuse_flow_fwd
  configs      BASELINE.json configs[2..4]: C3 training step (DDP), C4 SSMR windows, C5 4K x 31 timesteps (tools/bench_configs.py)
  cpu_baseline the reference's torch-op path (oracle/torch_oracle.py, kind "port") on the host cores, rank 0, bounded sample
  reference_gpu  the reference's torch-op path on this GPU (cuDNN sampler on / off), same workload
--impl reference times the CPU path as an arm of its own, on the same 16-pair workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IN, W_IN = 1080, 1920
H, W, PAIRS, NT = 1088, 1920, 16, 7
NPX = H * W
WORKLOAD = "ssm_original_1080p_b16_n7"
METRIC = "interpolated_1080p_frames_per_sec"


def config_dict(world):
    """`config` of the JSON line, identical in both arms"""
    return {"workload": WORKLOAD, "pairs_per_gpu": PAIRS, "timesteps": NT, "height": H, "width": W,
            "frames": "uint8 %dx%d images centred in %dx%d, normalised with the reference's mean/std" % (H_IN, W_IN, H, W),
            "frames_per_step": PAIRS * NT * world,
            "l2": "inputs_exceed_l2 (8.7 GB read, 18.9 GB written per step)",
            "coord_mode": "cpu (IEEE division, bit-matches the CPU reference)",
            "parallelism": "pairs sharded over %d rank(s), no collective" % world}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """DRAM bytes per launch of every profiled kernel, from the committed ncu --set full captures."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons of one GPU, polled DURING the timed region.
    NVML in a background thread (first sample within a millisecond of start(), one every 10 ms: a 0.3 s
    timed region still gets ~30 samples); `nvidia-smi -lms` as the fallback when pynvml is missing."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index, period_s=0.01):
        self.gpu = gpu_index
        self.period = period_s
        self.proc = self.thread = self.nvml = self.handle = None
        self.rows = []          # (sm_mhz, sm_max_mhz, power_w, reasons bitmask)
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                visible = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((sm, self.smax, power, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nvml is not None:
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        sm, smax, power, reasons = [], [], [], set()
        source = None
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            n = self.nvml
            bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
            for a, b, c, mask in self.rows:
                sm.append(a); smax.append(b); power.append(c)
                reasons.update(k for k, v in bits.items() if mask & v)
            source = "nvml"
        elif self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
                except ValueError:
                    continue
                reasons.update(n for n, v in zip(self.NAMES, parts[4:8]) if v == "Active")
            source = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": source}


# ---------------------------------------------------------------------------------------------
# The reference's own path: uint8 images -> load_batch padding + normalize_tensor -> compute_inputs +
# compute_output_image once per timestep, as its loops do (evaluate_interpolation_results.py:234-242).
def reference_sample(pairs, device="cpu", seed=42):
    from ssm_b200 import synthetic
    images = synthetic.images_u8(2 * pairs, H_IN, W_IN, seed=seed, device=device)
    flow4 = synthetic.flows(pairs, H, W, 4, flow_px=20.0, seed=seed + 1, device=device)
    out5 = synthetic.unet_out5(pairs, NT, H, W, seed=seed + 2, device=device)
    t = synthetic.timesteps(pairs, NT, device=device)
    return images, flow4, out5, t


def reference_step(sample, chunk=2):
    """the whole sample, `chunk` pairs at a time (the reference's op sequence materialises ~750 B/px of temporaries
    per call: 16 pairs at once would need ~25 GB)"""
    from oracle import torch_oracle
    images, flow4, out5, t = sample
    pairs = flow4.shape[0]
    count = 0
    with torch.no_grad():
        for b in range(0, pairs, chunk):
            e = min(b + chunk, pairs)
            # images are RGB here; load_batch_and_normalize takes cv2's BGR order
            frames = torch_oracle.load_batch_and_normalize(images[2 * b:2 * e].flip(-1).cpu().numpy(), device=flow4.device)[0]
            img6 = frames.reshape(e - b, 6, H, W)
            for n in range(NT):
                tn = t[b:e, n].view(-1, 1, 1, 1)
                in16 = torch_oracle.compute_inputs(img6, flow4[b:e], tn)
                frame = torch_oracle.compute_output_image(img6, in16, out5[b:e, n], tn)      # noqa: F841 (the result)
                count += frame.shape[0]
                del in16
    return count


def time_cpu_reference(steps, warmup, pairs):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = reference_sample(pairs)
    warm = reference_sample(1) if pairs > 1 else sample     # warm-up on one pair: it is untimed
    for _ in range(warmup):
        reference_step(warm)
    times = []
    budget = float(os.environ.get("SSM_REF_ARM_BUDGET_S", "420"))      # a slow host must not run into the driver's limit
    for _ in range(steps):
        t0 = time.perf_counter()
        reference_step(sample)
        times.append(time.perf_counter() - t0)
        if sum(times) > budget and len(times) >= 3:
            break
    return pairs * NT, times, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # the full 16-pair workload per step (about 9 s on 16 cores); warm-up steps run one pair
    frames, times, cores = time_cpu_reference(args.steps, min(args.warmup, 2), PAIRS)
    total = sum(times)
    value = frames * len(times) / total
    sample = ("the whole %d-pair x %d-timestep workload per step (%d frames) at %dx%d; uint8 frames normalised by the "
              "reference's expression, then compute_inputs + compute_output_image per timestep; torch CPU ops, %d threads; "
              "warm-up steps on one pair") % (PAIRS, NT, frames, H, W, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "steps_timed": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(1),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def time_reference_on_gpu(dev, chunk=4):
    """The reference's torch-op path on this GPU (what SURVEY.md section 2b names as the bar on the same box,
    scripts/models/layers.py:119): 16 pairs x 7 timesteps, in chunks of `chunk` pairs (the op sequence materialises
    ~750 B/px of temporaries per call), cuDNN's spatial-transformer sampler on and off (= ATen's grid_sampler)."""
    images, flow4, out5, t = reference_sample(PAIRS, device=dev)
    res = {}
    for name, flag in (("cudnn_on", True), ("cudnn_off", False)):
        prev = torch.backends.cudnn.enabled
        torch.backends.cudnn.enabled = flag
        try:
            def run():
                reference_step((images, flow4, out5, t), chunk=chunk)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            res[name] = {"ms_per_step": ms, "frames_per_s": PAIRS * NT / (ms * 1e-3)}
        except Exception as e:
            res[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        finally:
            torch.backends.cudnn.enabled = prev
        torch.cuda.empty_cache()
    res["note"] = ("oracle/torch_oracle.py == the reference's op sequence (~110 ATen launches and two host-built sampling grids "
                   "per timestep), run on cuda:0 outside the product; %d pairs x %d timesteps in chunks of %d pairs" % (PAIRS, NT, chunk))
    return res


# ---------------------------------------------------------------------------------------------
def whole_model_step(timeout_s=300):
    """tools/pipeline_step.py in a child process: uint8 host frames -> both U-Nets (stock torch/cuDNN, random init,
    channels-last + bf16 autocast; out of scope) + the path -> uint8 host frames, 2 pairs x 7 timesteps at 1080p.
    Reported next to the headline so that the share of the path in a whole step is visible; never the headline."""
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "pipeline_step.py")
    try:
        out = subprocess.run([sys.executable, tool, "--amp", "--channels-last", "--steps", "3", "--warmup", "2"],
                             capture_output=True, text=True, timeout=timeout_s)
        d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        keep = ("pairs", "timesteps", "height", "width", "amp_bf16_unets", "channels_last_unets", "ms_per_step",
                "frames_per_s", "h2d_bytes_per_step", "d2h_bytes_per_step", "path_kernels_ms", "path_ms_total",
                "path_share_of_step")
        res = {k: d[k] for k in keep if k in d}
        res["note"] = ("whole inference step incl. both U-Nets (stock torch/cuDNN, random init, out of scope): uint8 host "
                       "frames in, uint8 host frames out; context for the headline, not a bench value")
        return res
    except Exception as e:       # context only: never fail the bench line over it
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}


def host_copy_ceiling(dev, world, h2d_bytes, d2h_bytes, reps=3):
    """Plain pinned-memory copies of the SAME byte volumes one e2e step moves (h2d_bytes up, d2h_bytes down), the two
    directions concurrently on two streams, all ranks at once: the time the host side of this box needs for the step's
    transfers alone, i.e. the ceiling of any host-buffer entry point with these buffers."""
    import torch.distributed as dist
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    both()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        both()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        tt = torch.tensor([dt], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
    return {"ms_per_step_copies_only": 1e3 * dt, "h2d_gbs_per_gpu": h2d_bytes / dt / 1e9, "d2h_gbs_per_gpu": d2h_bytes / dt / 1e9,
            "aggregate_gbs_both_directions": (h2d_bytes + d2h_bytes) * world / dt / 1e9,
            "how": "the step's own byte volumes (%.2f GB up, %.2f GB down per rank) as plain pinned copies, both directions "
                   "concurrently on two streams, %d rank(s) at once, max time over ranks" % (h2d_bytes / 1e9, d2h_bytes / 1e9, world)}


PARTIAL = {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the backward-kernel timings")
    ap.add_argument("--no-variants", action="store_true", help="skip the smooth-flow, bf16-storage and layout variants")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[2..4] (C3 training step, C4 SSMR, C5 4K)")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference's torch-op path on this GPU")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    import ssm_b200
    from ssm_b200 import q8, sharding, synthetic

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the synthesis path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: 16 pairs per GPU; the partition function gives each rank whole pairs
    work = sharding.shard_work(PAIRS * world, NT, rank, world)
    B = len(work)
    assert B == PAIRS and all(t0 == 0 and t1 == NT for _, t0, t1 in work)
    seed0 = 42 + 1000 * rank
    images = synthetic.images_u8(2 * B, H_IN, W_IN, seed=seed0, device=dev)          # 2B x 1080 x 1920 x 3, RGB bytes
    lut = ssm_b200.normalisation_lut(device="cpu").to(dev)       # the CPU bit pattern of the reference's expression
    planar, quads, norm, (top, left) = q8.prepare(images, order="rgb", lut=lut, pad_values=lut[:, 0].tolist())
    img6 = planar.view(B, 6, H, W)                                 # the normalised frames (pre-step, as in the reference)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=seed0 + 1, device=dev)
    out5 = synthetic.unet_out5(B, NT, H, W, seed=seed0 + 2, device=dev)
    t = synthetic.timesteps(B, NT, device=dev)

    # caller-owned result buffers (out=): the steady-state loop makes no allocator calls, so no
    # cudaMalloc (which synchronises and maps pages) can land inside the timed region
    in16_buf = torch.empty((B, NT, 16, H, W), dtype=torch.float32, device=dev)
    frames_buf = torch.empty((B, NT, 3, H, W), dtype=torch.float32, device=dev)
    rgbx_buf = torch.empty((B, 2, H, W, 4), dtype=torch.float32, device=dev)
    L = ssm_b200._abi.lib()
    import ctypes
    quads_ptr, images_ptr = ctypes.c_void_p(quads.data_ptr()), ctypes.c_void_p(images.data_ptr())
    stream = ssm_b200._abi.stream_ptr(dev)

    def stage_quads():
        rc = L.ssm_quads_from_u8(images_ptr, images.stride(0), images.stride(1), 0, 2 * B, H_IN, W_IN, H, W, top, left, quads_ptr, stream)
        ssm_b200._abi.check(rc, "ssm_quads_from_u8")

    def step(flow=None, y=None, events=None):
        """One pass of the hot path over the batch: entry-table staging + the two fused launches."""
        f = flow4 if flow is None else flow
        with torch.no_grad():
            if events:
                events[0].record()
            stage_quads()
            if events:
                events[1].record()
            # img6=None: the six pass-through channels are looked up from the tables' own bytes (bit-identical to reading
            # the planar normalised frames, which this step then never touches)
            in16 = q8.flow_pack(None, quads, f, t, norm, n_timesteps=NT, out=in16_buf, lut=lut)
            if events:
                events[2].record()
            frames = q8.fuse_from_flow(quads, f, out5 if y is None else y, t, norm, out=frames_buf)
            if events:
                events[3].record()
        return in16, frames

    def step_fp32(flow=None, y=None, events=None):
        """The same step for normalised fp32 frames (the reference's tensor interface; round-1 headline)."""
        f = flow4 if flow is None else flow
        with torch.no_grad():
            if events:
                events[0].record()
            rgbx = ssm_b200.pack_frames(img6, out=rgbx_buf)
            if events:
                events[1].record()
            in16 = ssm_b200.flow_pack(img6, f, t, n_timesteps=NT, packed=rgbx, out=in16_buf)
            if events:
                events[2].record()
            frames = ssm_b200.fuse_from_flow(img6, f, out5 if y is None else y, t, packed=rgbx, out=frames_buf)
            if events:
                events[3].record()
        return in16, frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    if not args.no_clocks:
        sampler.start()
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(K):
        step(events=ev[k])
    end.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(end)
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = tt.item()
    frames_per_step = B * NT * world
    value = frames_per_step * K / (elapsed_ms * 1e-3)

    # ---- roofline: algorithmic bytes per SURVEY 8(d), the reference's fp32 tensors read / written once ----------------
    #   compute_inputs          (10 + 16 N) * 4 B/px per pair
    #   compute_output_image    (10 + 8 N) * 4 B/px per pair (estimated flows recomputed from the stage-1 flows)
    peak, peak_src = _peaks()
    traffic = _traffic()
    pack_bytes = (10 + 16 * NT) * 4 * NPX * B
    fuse_bytes = (10 + 8 * NT) * 4 * NPX * B

    def kernel_table(evs, names, staging_bytes):
        ms = [statistics.mean(e[i].elapsed_time(e[i + 1]) for e in evs) for i in range(3)]
        gbs = [staging_bytes / (ms[0] * 1e-3) / 1e9, pack_bytes / (ms[1] * 1e-3) / 1e9, fuse_bytes / (ms[2] * 1e-3) / 1e9]
        path_ms = sum(ms)
        path_gbs = (pack_bytes + fuse_bytes) / (path_ms * 1e-3) / 1e9
        step_ms = [e[0].elapsed_time(e[3]) for e in evs]
        tab = {
            names[0]: {"ms": ms[0], "moved_gbs": gbs[0], "note": "staging, overhead: charged to the path's time, not to its bytes"},
            names[1]: {"ms": ms[1], "algorithmic_gbs": gbs[1], "frac_of_peak": gbs[1] / peak},
            names[2]: {"ms": ms[2], "algorithmic_gbs": gbs[2], "frac_of_peak": gbs[2] / peak},
            "path": {"ms": path_ms, "algorithmic_gbs": path_gbs, "frac_of_peak": path_gbs / peak,
                     "frac_of_nominal_8tbs": path_gbs / 8000.0},
            "step_ms_distribution": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
        }
        return tab, ms, gbs

    quads_bytes = (2 * 3 * H_IN * W_IN + 2 * 16 * (H + 1) * (W + 1)) * B
    kernels, kms, kgbs = kernel_table(ev, ("quads_from_u8", "flow_pack_fwd_q8", "fuse_fwd_q8"), quads_bytes)
    roofline = {"bound": "hbm", "kernel": "flow_pack_fwd_q8_kernel", "achieved": kgbs[1], "peak": peak,
                "unit": "GB/s", "frac": kgbs[1] / peak, "traffic": traffic.get("flow_pack_fwd_q8_dram_bytes_per_launch"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": pack_bytes, "launch_ms": kms[1],
                "frac_of_nominal_8tbs": kgbs[1] / 8000.0}
    roofline_kernels = [
        {"kernel": "flow_pack_fwd_q8_kernel", "frames": "uint8", "bound": "hbm", "achieved": kgbs[1], "peak": peak, "unit": "GB/s",
         "frac": kgbs[1] / peak, "algorithmic_bytes_per_launch": pack_bytes, "launch_ms": kms[1],
         "traffic": traffic.get("flow_pack_fwd_q8_dram_bytes_per_launch")},
        {"kernel": "fuse_fwd_q8_kernel", "frames": "uint8", "bound": "hbm", "achieved": kgbs[2], "peak": peak, "unit": "GB/s",
         "frac": kgbs[2] / peak, "algorithmic_bytes_per_launch": fuse_bytes, "launch_ms": kms[2],
         "traffic": traffic.get("fuse_fwd_q8_dram_bytes_per_launch")},
    ]

    # safety net: if a SECONDARY section below fails (single-GPU runs), the headline measured above is still printed
    PARTIAL["line"] = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(world), "roofline": roofline, "roofline_kernels": list(roofline_kernels),
        "kernels": kernels, "gpu_launches": 3 * K, "clocks": clocks}

    def timed(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        es[0].record()
        for _ in range(reps):
            fn()
        es[1].record()
        torch.cuda.synchronize()
        return es[0].elapsed_time(es[1]) / reps

    def variant(ms, nbytes):
        return {"ms_per_step": ms, "frames_per_s": B * NT / (ms * 1e-3),
                "path_algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9, "path_frac_of_peak": nbytes / (ms * 1e-3) / 1e9 / peak}

    # ---- the same step on normalised fp32 frames (RGBx staging copy + fp32 gathers): every rank, 10 steps -------------
    for _ in range(3):
        step_fp32()
    torch.cuda.synchronize()
    ev32 = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(10)]
    for k in range(10):
        step_fp32(events=ev32[k])
    torch.cuda.synchronize()
    fp32_path, fms, fgbs = kernel_table(ev32, ("pack_frames", "flow_pack_fwd", "fuse_fwd"), (6 + 8) * 4 * NPX * B)
    fp32_path["frames_per_s_this_rank"] = B * NT / (fp32_path["path"]["ms"] * 1e-3)
    fp32_path["note"] = ("callers that hold normalised fp32 frames (the reference's tensor interface): RGBx staging copy + fp32 "
                         "16-byte gathers; round-1 headline workload, same flows / U-Net output / times")
    roofline_kernels += [
        {"kernel": "flow_pack_fwd_kernel<float>", "frames": "fp32", "bound": "hbm", "achieved": fgbs[1], "peak": peak, "unit": "GB/s",
         "frac": fgbs[1] / peak, "algorithmic_bytes_per_launch": pack_bytes, "launch_ms": fms[1],
         "traffic": traffic.get("flow_pack_fwd_dram_bytes_per_launch")},
        {"kernel": "fuse_fwd_kernel<float, RECOMP>", "frames": "fp32", "bound": "hbm", "achieved": fgbs[2], "peak": peak, "unit": "GB/s",
         "frac": fgbs[2] / peak, "algorithmic_bytes_per_launch": fuse_bytes, "launch_ms": fms[2],
         "traffic": traffic.get("fuse_fwd_dram_bytes_per_launch")},
    ]

    smooth = smooth_res = bf16 = layouts = None
    if rank == 0 and not args.no_variants:
        gen = torch.Generator(device=dev).manual_seed(7)

        def lowres(ch, div, scale, lead=()):
            c = torch.randn(lead + (B, ch, H // div, W // div), device=dev, generator=gen) * scale
            return torch.nn.functional.interpolate(c.view(-1, ch, H // div, W // div), size=(H, W), mode="bilinear",
                                                   align_corners=False).view(lead + (B, ch, H, W)).contiguous()
        # same kernels on a smooth flow field (control grid at 1/64 resolution): real optical flow is piecewise smooth;
        # the headline workload uses SURVEY 8(d)'s much rougher 1/8-resolution field
        flow_s = lowres(4, 64, 20.0)
        smooth = variant(timed(lambda: step(flow_s)), pack_bytes + fuse_bytes)
        smooth["fp32_frames_ms_per_step"] = timed(lambda: step_fp32(flow_s))
        # ... and with a spatially smooth stage-2 output as well (a trained U-Net's residual flows and
        # visibility logits are smooth; the surrogate of the headline workload is white noise)
        y_s = torch.empty_like(out5)
        for n in range(NT):
            y_s[:, n] = lowres(5, 16, 1.0)
        y_s[:, :, 0] *= 2.0
        y_s[:, :, 1:] *= 0.5
        smooth_res = variant(timed(lambda: step(flow_s, y_s)), pack_bytes + fuse_bytes)
        smooth_res["fp32_frames_ms_per_step"] = timed(lambda: step_fp32(flow_s, y_s))
        del flow_s, y_s
        # bf16 storage of every tensor, fp32 arithmetic (north_star tolerance 2e-2): half the bytes per element
        img_h, flow_h, out5_h = img6.bfloat16(), flow4.bfloat16(), out5.bfloat16()

        def step_bf16():
            with torch.no_grad():
                r = ssm_b200.pack_frames(img_h)
                a = ssm_b200.flow_pack(img_h, flow_h, t, n_timesteps=NT, packed=r)
                return a, ssm_b200.fuse_from_flow(img_h, flow_h, out5_h, t, packed=r)
        bf16_rgbx = variant(timed(step_bf16), (pack_bytes + fuse_bytes) // 2)
        in16_h = torch.empty((B, NT, 16, H, W), dtype=torch.bfloat16, device=dev)
        frames_h = torch.empty((B, NT, 3, H, W), dtype=torch.bfloat16, device=dev)

        def step_bf16_q8():
            with torch.no_grad():
                stage_quads()
                a = q8.flow_pack(img_h, quads, flow_h, t, norm, n_timesteps=NT, out=in16_h)
                return a, q8.fuse_from_flow(quads, flow_h, out5_h, t, norm, out=frames_h)
        bf16 = variant(timed(step_bf16_q8), (pack_bytes + fuse_bytes) // 2)
        bf16["note"] = ("bf16 storage of every tensor (planar frames, flows, U-Net output, both results), fp32 arithmetic, the warps "
                        "gathering from the uint8 entry tables as in the headline; not the headline (the reference is fp32). Half the "
                        "bytes for the same instruction count: the kernels are bound by instruction issue / the FMA pipe, not HBM "
                        "(DESIGN 5.1b).  bf16_rgbx_frames: the round-1 form (frames as a bf16 RGBx copy, 8-byte gathers)")
        bf16["bf16_rgbx_frames"] = bf16_rgbx
        del img_h, flow_h, in16_h, frames_h
        # the layouts either side of a channels-last stage-2 U-Net under bf16 autocast:
        # compute_inputs writes B x N x H x W x 16 bf16, compute_output_image reads a bf16 U-Net output
        nhwc_buf = torch.empty((B, NT, H, W, 16), dtype=torch.bfloat16, device=dev).permute(0, 1, 4, 2, 3)

        def step_layouts():
            with torch.no_grad():
                stage_quads()
                a = q8.flow_pack(img6, quads, flow4, t, norm, n_timesteps=NT, out=nhwc_buf, channels_last_dtype=torch.bfloat16)
                return a, q8.fuse_from_flow(quads, flow4, out5_h, t, norm, out=frames_buf)
        layouts = variant(timed(step_layouts), ((10 * 4 + 16 * 2 * NT) + (10 * 4 + (5 * 2 + 3 * 4) * NT)) * NPX * B)
        layouts["note"] = ("compute_inputs written channels-last bf16 (what conv1a consumes under channels-last bf16 autocast) and the "
                           "bf16 U-Net output read directly; uint8 frames, fp32 flows and result; not the headline")
        del out5_h, nhwc_buf

    # ---- training backward of the fp32 kernels (flow / U-Net-output gradients; frames are data) ----
    train = None
    if rank == 0 and not args.no_train:
        fg, yg = flow4.clone().requires_grad_(True), out5.clone().requires_grad_(True)
        rgbx = ssm_b200.pack_frames(img6)
        in16 = ssm_b200.flow_pack(img6, fg, t, n_timesteps=NT, packed=rgbx)
        frames = ssm_b200.fuse_from_flow(img6, fg, yg, t, packed=rgbx)
        g3, g16 = torch.randn_like(frames), torch.randn_like(in16)
        tb = {"fuse_flow_bwd": [], "flow_pack_bwd": []}
        for i in range(8):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            torch.autograd.grad(frames, (fg, yg), g3, retain_graph=True)
            e[1].record()
            torch.autograd.grad(in16, (fg,), g16, retain_graph=True)
            e[2].record()
            torch.cuda.synchronize()
            if i >= 3:
                tb["fuse_flow_bwd"].append(e[0].elapsed_time(e[1]))
                tb["flow_pack_bwd"].append(e[1].elapsed_time(e[2]))
        nbytes = {"fuse_flow_bwd": (10 + 13 * NT) * 4 * NPX * B, "flow_pack_bwd": (14 + 10 * NT) * 4 * NPX * B}
        train = {}
        for k, v in tb.items():
            ms = statistics.median(v)
            train[k] = {"ms": ms, "algorithmic_gbs": nbytes[k] / (ms * 1e-3) / 1e9,
                        "frac_of_peak": nbytes[k] / (ms * 1e-3) / 1e9 / peak}
            roofline_kernels.append({"kernel": k + " (fp32 frames)", "frames": "fp32", "bound": "hbm", "achieved": train[k]["algorithmic_gbs"],
                                     "peak": peak, "unit": "GB/s", "frac": train[k]["frac_of_peak"],
                                     "algorithmic_bytes_per_launch": nbytes[k], "launch_ms": ms,
                                     "traffic": traffic.get(k + "_dram_bytes_per_launch")})
        # the same backward with image gradients (frames require grad): deterministic fixed-point scatter on top
        ig = img6.clone().requires_grad_(True)
        frames_i = ssm_b200.fuse_from_flow(ig, fg, yg, t, packed=rgbx)
        ims = []
        for i in range(4):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            torch.autograd.grad(frames_i, (ig, fg, yg), g3, retain_graph=True)
            e[1].record()
            torch.cuda.synchronize()
            if i >= 1:
                ims.append(e[0].elapsed_time(e[1]))
        ms = statistics.median(ims)
        nb = nbytes["fuse_flow_bwd"] + 6 * 4 * NPX * B
        train["fuse_flow_bwd_with_image_grad"] = {"ms": ms, "algorithmic_gbs": nb / (ms * 1e-3) / 1e9, "frac_of_peak": nb / (ms * 1e-3) / 1e9 / peak,
                                                  "times_the_gather_only_backward": ms / train["fuse_flow_bwd"]["ms"]}
        del fg, yg, rgbx, in16, frames, g3, g16, ig, frames_i
        torch.cuda.empty_cache()
        # stand-alone warp (a1, layers.py:73-120) on the 16 first frames: planar gathers and the RGBx copy
        xw, fw = img6[:, 0:3].contiguous().requires_grad_(True), flow4[:, 0:2].contiguous().requires_grad_(True)
        gw = torch.randn_like(xw)
        pk = ssm_b200.pack_image(xw)
        warp_t = {}
        for name, kw in (("planar", {}), ("rgbx", {"packed": pk})):
            warp_t["warp_fwd_" + name] = (timed(lambda: ssm_b200.warp(xw.detach(), fw.detach(), **kw)), 8)
            yw = ssm_b200.warp(xw.detach(), fw, **kw)                    # frames are data: flow gradient only (a gather)
            warp_t["warp_bwd_flow_" + name] = (timed(lambda: torch.autograd.grad(yw, (fw,), gw, retain_graph=True)), 10)
            del yw
        yw = ssm_b200.warp(xw, fw)                                        # + image gradient (segmented scatter)
        warp_t["warp_bwd_flow_and_image"] = (timed(lambda: torch.autograd.grad(yw, (xw, fw), gw, retain_graph=True), reps=3, warm=1), 13)
        del yw
        warp_t["pack_image"] = (timed(lambda: ssm_b200.pack_image(xw, out=pk)), 7)
        for k, (ms, per_px) in warp_t.items():
            nb = per_px * 4 * NPX * B
            train[k] = {"ms": ms, "algorithmic_gbs": nb / (ms * 1e-3) / 1e9, "frac_of_peak": nb / (ms * 1e-3) / 1e9 / peak}
        del xw, fw, gw, pk
    del in16_buf, rgbx_buf
    torch.cuda.empty_cache()

    # ---- e2e: host buffers through the C-ABI host entry points --------------------------------
    e2e = e2e_fp32 = ceiling = None
    if not args.no_e2e:
        # pinned buffers on the NUMA node the GPU hangs off (each rank binds to its own GPU's node)
        with sharding.numa_local_to_gpu(local_rank) as placement:
            h_images = images.view(B, 2, H_IN, W_IN, 3).cpu().pin_memory()
            h_flow, h_out5 = flow4.cpu().pin_memory(), out5.bfloat16().cpu().pin_memory()
            h_t = t.cpu()
            h_out = torch.empty((B, NT, H_IN, W_IN, 3), dtype=torch.uint8, pin_memory=True)
        h2d = h_images.numel() + h_flow.numel() * 4 + h_out5.numel() * 2 + h_t.numel() * 4
        d2h = h_out.numel()
        ceiling = host_copy_ceiling(dev, world, h2d, d2h)
        scratch = torch.empty(q8.synthesize_host_scratch_bytes(B, NT, H_IN, W_IN, torch.bfloat16), dtype=torch.uint8, device=dev)
        q8.synthesize_host(h_images, h_flow, h_out5, h_t, order="rgb", out=h_out, scratch=scratch)          # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = q8.synthesize_host(h_images, h_flow, h_out5, h_t, order="rgb", out=h_out, scratch=scratch)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = tt.item()
        assert d2h == res.numel()
        e2e_value = frames_per_step * args.e2e_steps / dt
        bound = frames_per_step / (ceiling["ms_per_step_copies_only"] * 1e-3)
        e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": "ssm_synthesize_host_u8 (pinned host buffers: uint8 frames, fp32 stage-1 flows, bf16 U-Net output in; uint8 "
                      "interpolated images out; 3-slot copy/compute pipeline)",
               "host_placement": placement.info, "host_copy_ceiling": ceiling,
               "host_copy_ceiling_frames_per_s": bound, "frac_of_host_copy_ceiling": e2e_value / bound}
        del h_images, h_out5, res, h_out, scratch
        torch.cuda.empty_cache()
        # the fp32 host entry (round-1 e2e): fp32 frames, flows and U-Net output in, fp32 frames out
        if rank == 0 and world == 1:
            h_img, h_out5f = img6.cpu().pin_memory(), out5.cpu().pin_memory()
            h_out3 = torch.empty((B, NT, 3, H, W), dtype=torch.float32, pin_memory=True)
            scratch = torch.empty(ssm_b200.synthesize_host_scratch_bytes(B, NT, H, W), dtype=torch.uint8, device=dev)
            ssm_b200.synthesize_host(h_img, h_flow, h_out5f, h_t, out=h_out3, scratch=scratch)
            t0 = time.perf_counter()
            for _ in range(2):
                ssm_b200.synthesize_host(h_img, h_flow, h_out5f, h_t, out=h_out3, scratch=scratch)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 2
            e2e_fp32 = {"value": B * NT / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt,
                        "h2d_bytes_per_step": (h_img.numel() + h_flow.numel() + h_out5f.numel() + h_t.numel()) * 4,
                        "d2h_bytes_per_step": h_out3.numel() * 4, "api": "ssm_synthesize_host (fp32 planes in and out; round-1 e2e)"}
            del h_img, h_out5f, h_out3, scratch
        del h_flow
        torch.cuda.empty_cache()

    # ---- CPU baseline: the reference's torch-op path on the host cores, rank 0, N=1 only ------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        frames_c, times, cores = time_cpu_reference(steps=3, warmup=1, pairs=1)
        best = min(times)
        cpu_baseline = {"value": frames_c / best, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": "1 pair x %d timesteps at %dx%d (uint8 frames normalised by the reference's expression, then "
                                  "compute_inputs + compute_output_image per timestep), best of 3 after 1 warm-up; torch CPU ops "
                                  "(oracle/torch_oracle.py == the reference's op sequence)" % (NT, H, W),
                        "ms_per_frame": 1e3 * best / frames_c}

    # ---- the reference's torch-op path on this GPU (cuDNN sampler on / off): the same-box bar ------------------------
    reference_gpu = None
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        del frames_buf
        torch.cuda.empty_cache()
        reference_gpu = time_reference_on_gpu(dev)
        reference_gpu["speedup_of_this_path"] = {k: (B * NT / (kernels["path"]["ms"] * 1e-3)) / v["frames_per_s"]
                                                 for k, v in reference_gpu.items() if isinstance(v, dict) and "frames_per_s" in v}

    # ---- BASELINE configs[2..4] on all ranks --------------------------------------------------------------------------
    configs = None
    if not args.no_configs:
        del images, planar, img6, quads, flow4, out5
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        configs = bench_configs.run_all(world, rank, dev, peak_gbs=peak)

    # ---- context, not the headline: where the path sits in a whole inference step with the two U-Nets ------------
    whole_model = None
    if rank == 0 and world == 1 and not args.no_variants:
        whole_model = whole_model_step()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world),
            "roofline": roofline, "roofline_kernels": roofline_kernels, "kernels": kernels, "fp32_frames_path": fp32_path,
            "train_kernels": train, "smooth_flow_variant": smooth,
            "smooth_flow_and_unet_output_variant": smooth_res, "bf16_storage_variant": bf16, "unet_layouts_variant": layouts,
            "configs": configs, "reference_gpu": reference_gpu,
            "whole_model_context": whole_model, "cpu_baseline": cpu_baseline,
            "e2e": e2e, "e2e_fp32_buffers": e2e_fp32, "gpu_launches": 3 * K, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except Exception as exc:       # a secondary section failed after the headline was measured: report both
        if PARTIAL.get("line") and int(os.environ.get("RANK", "0")) == 0 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            import traceback
            traceback.print_exc()
            line = PARTIAL["line"]
            line["secondary_section_error"] = "%s: %s" % (type(exc).__name__, str(exc)[:300])
            print(json.dumps(line))
            sys.exit(0)
        raise
