#!/usr/bin/env python
"""bench.py -- throughput of the intermediate-frame synthesis path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 16 synthetic 1088x1920 (1080p padded to /32) frame pairs per
GPU, 7 intermediate timesteps, fp32.  One step = one pass of the hot path over the batch:
an RGBx staging copy of the frames (one launch), compute_inputs for all 7 timesteps (one fused launch)
and extract_outputs/compute_output_image for all 7 timesteps (one fused launch, estimated flows
recomputed in-kernel from the stage-1 flows: ssm_fuse_flow_fwd) = 112 interpolated frames per GPU.  The two flow U-Nets are out of
scope (they stay on PyTorch/cuDNN); the stage-2 output is a seeded surrogate.

  value      frames/s, inputs resident in HBM, device-timed with CUDA events, max over ranks
  e2e        same metric through ssm_synthesize_host: pinned HOST buffers in, fused frames out,
             host<->device copies inside the timed region
  roofline   dominant kernel (flow_pack_fwd) algorithmic bytes / event-timed duration vs measured HBM peak
  cpu_baseline  the reference's torch-op path (oracle/torch_oracle.py, kind "port") on the host cores,
             rank 0, bounded sample
--impl reference times that CPU path as an arm of its own.
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

H, W, PAIRS, NT = 1088, 1920, 16, 7
NPX = H * W
WORKLOAD = "ssm_original_1080p_b16_n7"
METRIC = "interpolated_1080p_frames_per_sec"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("flow_pack_fwd_dram_bytes_per_launch")
    except Exception:
        return None


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
def cpu_reference_step(sample):
    """One bounded sample of the reference CPU path: compute_inputs + compute_output_image once per
    timestep, as the reference's loops do (evaluate_interpolation_results.py:234-242)."""
    from oracle import torch_oracle
    img6, flow4, out5, t = sample
    outs = []
    with torch.no_grad():
        for n in range(t.shape[1]):
            tn = t[:, n].view(-1, 1, 1, 1)
            in16 = torch_oracle.compute_inputs(img6, flow4, tn)
            outs.append(torch_oracle.compute_output_image(img6, in16, out5[:, n], tn))
    return outs


def cpu_sample(pairs=1):
    from ssm_b200 import synthetic
    img6 = synthetic.frames(pairs, H, W, seed=42)
    flow4 = synthetic.flows(pairs, H, W, 4, flow_px=20.0, seed=43)
    out5 = synthetic.unet_out5(pairs, NT, H, W, seed=44)
    t = synthetic.timesteps(pairs, NT)
    return img6, flow4, out5, t


def time_cpu_reference(steps, warmup, pairs=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = cpu_sample(pairs)
    for _ in range(warmup):
        cpu_reference_step(sample)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(sample)
        times.append(time.perf_counter() - t0)
    frames = pairs * NT
    return frames, times, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    frames, times, cores = time_cpu_reference(args.steps, args.warmup)
    total = sum(times)
    value = frames * len(times) / total
    sample = "%d pair x %d timesteps at %dx%d per step (of the %d-pair workload), torch CPU ops, %d threads" % (
        1, NT, H, W, PAIRS, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": PAIRS, "timesteps": NT, "height": H, "width": W,
                   "frames_per_step": frames, "l2": "inputs_exceed_l2"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


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
    ap.add_argument("--no-variants", action="store_true", help="skip the smooth-flow and bf16-storage variants")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    import ssm_b200
    from ssm_b200 import sharding, synthetic

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
    img6 = synthetic.frames(B, H, W, seed=seed0, device=dev)
    flow4 = synthetic.flows(B, H, W, 4, flow_px=20.0, seed=seed0 + 1, device=dev)
    out5 = synthetic.unet_out5(B, NT, H, W, seed=seed0 + 2, device=dev)
    t = synthetic.timesteps(B, NT, device=dev)

    # caller-owned result buffers (out=): the steady-state loop makes no allocator calls, so no
    # cudaMalloc (which synchronises and maps pages) can land inside the timed region
    rgbx_buf = torch.empty((B, 2, H, W, 4), dtype=torch.float32, device=dev)
    in16_buf = torch.empty((B, NT, 16, H, W), dtype=torch.float32, device=dev)
    frames_buf = torch.empty((B, NT, 3, H, W), dtype=torch.float32, device=dev)

    def step(flow=None, y=None, events=None):
        """One pass of the hot path over the batch: RGBx staging copy + the two fused launches."""
        f = flow4 if flow is None else flow
        with torch.no_grad():
            if events:
                events[0].record()
            rgbx = ssm_b200.pack_frames(img6, out=rgbx_buf)      # RGBx staging copy, shared by both kernels
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
    rgbx_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    pack_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    fuse_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in ev)
    step_ms = [e[0].elapsed_time(e[3]) for e in ev]
    frames_per_step = B * NT * world
    value = frames_per_step * K / (elapsed_ms * 1e-3)

    # roofline of the dominant kernel: algorithmic bytes = (10 + 16 N) * 4 B/px per pair
    peak, peak_src = _peaks()
    pack_bytes = (10 + 16 * NT) * 4 * NPX * B
    # a3+a4 with the estimated flows recomputed from flow4: reads I (6) + F (4) per pair and out5 (5) per
    # timestep, writes 3 per timestep (the reference-shaped call that re-reads in16[:, 6:10] is (6 + 12 N))
    fuse_bytes = (10 + 8 * NT) * 4 * NPX * B
    pack_gbs = pack_bytes / (pack_ms * 1e-3) / 1e9
    fuse_gbs = fuse_bytes / (fuse_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "flow_pack_fwd_kernel<float>", "achieved": pack_gbs, "peak": peak,
                "unit": "GB/s", "frac": pack_gbs / peak, "traffic": _traffic(), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": pack_bytes, "launch_ms": pack_ms,
                "frac_of_nominal_8tbs": pack_gbs / 8000.0}
    # the RGBx staging copy is overhead, not algorithmic traffic: it is charged to the path's time
    # but not to its bytes
    rgbx_bytes = (6 + 8) * 4 * NPX * B
    path_ms = rgbx_ms + pack_ms + fuse_ms
    path_gbs = (pack_bytes + fuse_bytes) / (path_ms * 1e-3) / 1e9
    kernels = {
        "pack_frames": {"ms": rgbx_ms, "moved_gbs": rgbx_bytes / (rgbx_ms * 1e-3) / 1e9, "note": "staging copy, overhead"},
        "flow_pack_fwd": {"ms": pack_ms, "algorithmic_gbs": pack_gbs, "frac_of_peak": pack_gbs / peak},
        "fuse_fwd": {"ms": fuse_ms, "algorithmic_gbs": fuse_gbs, "frac_of_peak": fuse_gbs / peak},
        "path": {"ms": path_ms, "algorithmic_gbs": path_gbs, "frac_of_peak": path_gbs / peak},
        "step_ms_distribution": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
    }

    # same kernels on a smooth flow field (control grid at 1/64 resolution): real optical flow is
    # piecewise smooth; the headline workload above uses SURVEY 8(d)'s much rougher 1/8-resolution field
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

    smooth = smooth_res = bf16 = layouts = None
    if rank == 0 and not args.no_variants:
        gen = torch.Generator(device=dev).manual_seed(7)

        def lowres(ch, div, scale, lead=()):
            c = torch.randn(lead + (B, ch, H // div, W // div), device=dev, generator=gen) * scale
            return torch.nn.functional.interpolate(c.view(-1, ch, H // div, W // div), size=(H, W), mode="bilinear",
                                                   align_corners=False).view(lead + (B, ch, H, W)).contiguous()
        flow_s = lowres(4, 64, 20.0)
        smooth = variant(timed(lambda: step(flow_s)), pack_bytes + fuse_bytes)
        # ... and with a spatially smooth stage-2 output as well (a trained U-Net's residual flows and
        # visibility logits are smooth; the surrogate of the headline workload is white noise)
        y_s = torch.empty_like(out5)
        for n in range(NT):
            y_s[:, n] = lowres(5, 16, 1.0)
        y_s[:, :, 0] *= 2.0
        y_s[:, :, 1:] *= 0.5
        smooth_res = variant(timed(lambda: step(flow_s, y_s)), pack_bytes + fuse_bytes)
        del flow_s, y_s
        # bf16 storage, fp32 arithmetic (north_star tolerance 2e-2): half the bytes per element
        img_h, flow_h, out5_h = img6.bfloat16(), flow4.bfloat16(), out5.bfloat16()

        def step_bf16():
            with torch.no_grad():
                r = ssm_b200.pack_frames(img_h)
                a = ssm_b200.flow_pack(img_h, flow_h, t, n_timesteps=NT, packed=r)
                return a, ssm_b200.fuse_from_flow(img_h, flow_h, out5_h, t, packed=r)
        bf16 = variant(timed(step_bf16), (pack_bytes + fuse_bytes) // 2)
        bf16["note"] = "bf16 storage of every tensor, fp32 arithmetic; not the headline (the reference is fp32)"
        del img_h, flow_h
        # the layouts either side of a channels-last stage-2 U-Net under bf16 autocast:
        # compute_inputs writes B x N x H x W x 16 bf16, compute_output_image reads a bf16 U-Net output;
        # frames, flows and the fused frames stay fp32
        nhwc_buf = torch.empty((B, NT, H, W, 16), dtype=torch.bfloat16, device=dev).permute(0, 1, 4, 2, 3)

        def step_layouts():
            with torch.no_grad():
                r = ssm_b200.pack_frames(img6, out=rgbx_buf)
                a = ssm_b200.flow_pack_channels_last(img6, flow4, t, n_timesteps=NT, dtype=torch.bfloat16, packed=r,
                                                     out=nhwc_buf)
                return a, ssm_b200.fuse_from_flow(img6, flow4, out5_h, t, packed=r, out=frames_buf)
        layouts = variant(timed(step_layouts), ((10 * 4 + 16 * 2 * NT) + (10 * 4 + (5 * 2 + 3 * 4) * NT)) * NPX * B)

        # what the generic plumbing does for the same U-Net: planar fp32 compute_inputs, one torch pass that
        # converts it to channels-last bf16, one that widens the bf16 U-Net output to fp32
        out5_f = torch.empty_like(out5)

        def step_generic():
            with torch.no_grad():
                r = ssm_b200.pack_frames(img6, out=rgbx_buf)
                a = ssm_b200.flow_pack(img6, flow4, t, n_timesteps=NT, packed=r, out=in16_buf)
                nhwc_buf.copy_(a)
                out5_f.copy_(out5_h)
                return ssm_b200.fuse_from_flow(img6, flow4, out5_f, t, packed=r, out=frames_buf)
        layouts["generic_ms_per_step"] = timed(step_generic)
        layouts["note"] = ("compute_inputs written channels-last bf16 (what conv1a consumes under channels-last bf16 "
                           "autocast) and the bf16 U-Net output read directly; generic_ms_per_step = the planar fp32 "
                           "kernels plus the two torch conversion passes they need for the same U-Net; fp32 "
                           "frames/flows/result; not the headline")
        del out5_h, nhwc_buf, out5_f

    # ---- training backward of the same kernels (flow / U-Net-output gradients; frames are data) ----
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
        del fg, yg, rgbx, in16, frames, g3, g16

    # ---- e2e: host buffers through the C-ABI host entry point --------------------------------
    e2e = None
    if not args.no_e2e:
        # pinned buffers on the NUMA node the GPU hangs off (each rank binds to its own GPU's node)
        with sharding.numa_local_to_gpu(local_rank) as placement:
            h_img, h_flow, h_out5 = img6.cpu().pin_memory(), flow4.cpu().pin_memory(), out5.cpu().pin_memory()
            h_t = t.cpu()
            h_out = torch.empty((B, NT, 3, H, W), dtype=torch.float32, pin_memory=True)
        scratch = torch.empty(ssm_b200.synthesize_host_scratch_bytes(B, NT, H, W), dtype=torch.uint8, device=dev)
        ssm_b200.synthesize_host(h_img, h_flow, h_out5, h_t, out=h_out, scratch=scratch)          # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = ssm_b200.synthesize_host(h_img, h_flow, h_out5, h_t, out=h_out, scratch=scratch)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = tt.item()
        h2d = (h_img.numel() + h_flow.numel() + h_out5.numel() + h_t.numel()) * 4
        d2h = res.numel() * 4
        e2e = {"value": frames_per_step * args.e2e_steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": "ssm_synthesize_host (pinned host buffers, 3-slot copy/compute pipeline)",
               "host_placement": placement.info}
        del h_img, h_flow, h_out5, res, h_out, scratch

    # ---- CPU baseline: the reference's torch-op path on the host cores, rank 0, N=1 only ------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        frames_c, times, cores = time_cpu_reference(steps=3, warmup=1)
        best = min(times)
        cpu_baseline = {"value": frames_c / best, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": "1 pair x %d timesteps at %dx%d, best of 3 after 1 warm-up; torch CPU ops "
                                  "(oracle/torch_oracle.py == the reference's op sequence)" % (NT, H, W),
                        "ms_per_frame": 1e3 * best / frames_c}

    # ---- context, not the headline: where the path sits in a whole inference step with the two U-Nets ------------
    whole_model = None
    if rank == 0 and world == 1 and not args.no_variants:
        whole_model = whole_model_step()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": PAIRS, "timesteps": NT, "height": H, "width": W,
                       "frames_per_step": frames_per_step, "l2": "inputs_exceed_l2 (8.7 GB read, 18.9 GB written per step)",
                       "coord_mode": "cpu (IEEE division, bit-matches the CPU reference)",
                       "parallelism": "pairs sharded over %d rank(s), no collective" % world},
            "roofline": roofline, "kernels": kernels, "train_kernels": train, "smooth_flow_variant": smooth,
            "smooth_flow_and_unet_output_variant": smooth_res, "bf16_storage_variant": bf16, "unet_layouts_variant": layouts,
            "whole_model_context": whole_model, "cpu_baseline": cpu_baseline,
            "e2e": e2e, "gpu_launches": 3 * K, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
