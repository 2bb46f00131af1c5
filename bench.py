#!/usr/bin/env python
"""bench.py — guided-sampling frames/sec of the Climate2Weather hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] at N=1 — guided predictor-corrector sampling of a 1-week hourly
trajectory, L = 168 frames of 4 x 128 x 128, ScoreUNet of configs/sda_unet.yml (random init), Markov order k = 6
(156 windows), 256 denoising steps, 0 corrections, tau 0.5, exact_grad False, observation = every 6th frame 16x16
tile means (exp/configs/000_on-model-eval/s16_t6.yml).  For N > 1 the trajectory is time-sharded with 156 windows
per GPU (L = 12 + 156 N frames: weak scaling) and the k boundary frames are exchanged after every update.

A "step" is one denoising step over the whole trajectory: window score (156 UNet windows per GPU) + fused guidance
and predictor update (+ halo exchange).  `value` = frames / (256 * mean step time): frames per second of a full
256-step sampling run, state resident in HBM, timed with CUDA events over exactly K steps.  `e2e` = the same metric
through the public API (SDAPipeline.sample) with pinned HOST noise in, HOST result out and the per-step NaN-flag
read the reference does (src/thor/pipelines.py:90), all inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SAMPLER_STEPS = 256
K_ORDER = 6
C, H, W = 4, 128, 128
WINDOWS_PER_GPU = 156
T_STEP, S_STEP = 6, 16
STD = [0.1692666615037876, 0.0425178630338289, 0.3268027589410125, 0.3268027589410125]
GAMMA = 0.0007196856730011522
F_WIN_CONV = 115.134e9  # Conv2d FLOPs per window forward
F_WIN = 116.0e9
ARCH = dict(channels=52, embedding_dim=512, hidden_channels=[128, 128, 256, 384, 512], hidden_blocks=[3, 3, 3, 3, 3],
            attention_levels=[4])


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line).  The sampler is
    started BEFORE the warm-up (nvidia-smi needs a few hundred ms to come up) and every row is stamped on arrival;
    the summary uses the rows between mark_start() and mark_end()."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.05]
        rows = inside or [r for _, r in self.rows]
        sm, mx, reasons = [], 0, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "window": "timed region" if inside else "whole run"}


def make_problem(L: int, seed: int = 0):
    import torch

    import climate2weather_b200 as c2w

    torch.manual_seed(seed)
    net = c2w.ScoreUNet(**ARCH)
    g = torch.Generator().manual_seed(seed + 1)
    noise = torch.randn(L, C, H, W, generator=g)
    truth = torch.randn(L, C, H, W, generator=g)
    cg = c2w.CoarseGrain(T_STEP, S_STEP)
    y = cg(truth)
    return net, noise, y, cg


# ====================================================================================================== ours
def run_ours(args):
    import torch
    import torch.distributed as dist

    import climate2weather_b200 as c2w
    from climate2weather_b200 import _lib
    from climate2weather_b200.score import _mu_sigma

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = 2 * K_ORDER + WINDOWS_PER_GPU * world if args.frames is None else args.frames
    net, noise, y, cg = make_problem(L)
    net = net.to(dev)
    pipe = c2w.SDAPipeline()
    sf = c2w.BatchedScoreFunction(net, markov_order=K_ORDER, noise_process=pipe, batch_size=args.chunk, device=dev)
    sf.condition_on(A=cg, y=y, std=torch.tensor(STD).reshape(1, C, 1, 1), gamma=GAMMA, exact_grad=args.exact_grad)
    if world > 1:
        sf.enable_time_sharding()
    rt = sf.runtime(noise)
    group = None
    lib = _lib.load()
    times = torch.linspace(1, 0, SAMPLER_STEPS + 1)
    dt = 1 / SAMPLER_STEPS

    def one_step(i: int):
        t = times[i % SAMPLER_STEPS]
        mu, sigma = _mu_sigma(pipe, t)
        mu_n, sigma_n = _mu_sigma(pipe, t - dt)
        rt.score(float(t), group)
        rt.predictor(mu, sigma, mu_n, sigma_n)
        rt.halo(group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region: exactly K steps, CUDA events on the launch stream
    with ClockSampler(local_rank) as clocks:
        rt.load(noise)
        for i in range(args.warmup):
            one_step(i)
        rt.load(noise)  # restart the trajectory so the timed steps see the schedule from t = 1
        barrier()
        launches0 = lib.c2w_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.mark_start()
        e0.record()
        for i in range(args.steps):
            one_step(i)
        e1.record()
        barrier()
        clocks.mark_end()
    ms_total = e0.elapsed_time(e1)
    launches = lib.c2w_launch_count() - launches0
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = tmax.item() / args.steps
    value = L / (SAMPLER_STEPS * ms_step / 1e3)
    rt.check_finite()

    # ---- roofline pass: event-time every forward-pass launch of a few steps (same process, right after)
    eng = rt.engine
    _lib.check(lib.c2w_set_timing(eng.handle, 1), "c2w_set_timing")
    ms2 = (ctypes.c_double * 2)()
    n2 = (ctypes.c_int64 * 2)()
    nprof = 0 if args.profile else min(args.steps, 4)
    for i in range(nprof):
        one_step(i)
    _lib.check(lib.c2w_timing_read(eng.handle, ms2, n2), "c2w_timing_read")
    _lib.check(lib.c2w_set_timing(eng.handle, 0), "c2w_set_timing")
    n_win_local = rt.plan.win_hi - rt.plan.win_lo
    conv_ms_step = ms2[0] / max(nprof, 1)
    other_ms_step = ms2[1] / max(nprof, 1)
    peak_tf, peak_gbs, peak_src = measured_peaks()
    conv_tf = F_WIN_CONV * n_win_local / (conv_ms_step * 1e-3) / 1e12 if conv_ms_step > 0 else 0.0
    roofline = {
        "kernel": "conv_gemm_tcgen05_kernel (K1, all %d conv/GEMM launches of a step)" % (n2[0] // max(nprof, 1)),
        "bound": "tensor", "achieved": round(conv_tf, 1), "peak": peak_tf, "unit": "TFLOP/s",
        "frac": round(conv_tf / peak_tf, 4), "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": None,
        "algorithmic_flops_per_step": F_WIN_CONV * n_win_local,
        "avg_launch_ms": round(ms2[0] / max(1, n2[0]), 5), "k1_ms_per_step": round(conv_ms_step, 3),
        "other_fwd_kernels_ms_per_step": round(other_ms_step, 3),
        "k1_share_of_step": round(conv_ms_step / ms_step, 4),
    }

    # ---- end to end through the public API: pinned host noise in, host result out, per-step NaN-flag read
    e2e = None
    if not args.no_e2e:
        pipe2 = c2w.SDAPipeline()
        pipe2.nan_check_every = 1
        noise_pinned = noise.pin_memory()
        ke = max(1, args.e2e_steps)  # one full sample() call of this many denoising steps (independent of --steps)
        barrier()
        t0 = time.perf_counter()
        out = pipe2.sample(sf, noise_pinned, steps=ke, corrections=0, tau=0.5, show_progressbar=False)
        barrier()
        dt_e2e = time.perf_counter() - t0
        tm = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        assert out.device.type == "cpu" and bool(torch.isfinite(out).all())
        local_frames = rt.plan.n_local
        e2e = {"value": round(L / (SAMPLER_STEPS * tm.item() / ke), 3), "unit": "frames/s",
               "h2d_bytes_per_step": int(local_frames * C * H * W * 4 / ke),
               "d2h_bytes_per_step": int(rt.plan.own_n * C * H * W * 4 / ke) + 4,
               "api": f"SDAPipeline.sample(BatchedScoreFunction, pinned host noise, steps={ke}) -> host tensor",
               "steps": ke}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(sample_windows=args.cpu_windows, steps=1)

    if rank == 0:
        line = {
            "metric": "guided-sampling frames/sec", "value": round(value, 3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"config2: guided PC sampling, L={L} frames ({n_win_local} windows/GPU) of "
                                   f"{C}x{H}x{W}, sda_unet.yml ScoreUNet k={K_ORDER}, {SAMPLER_STEPS} steps, "
                                   f"0 corrections, {'exact-grad (UNet VJP)' if args.exact_grad else 'approx-grad'} guidance "
                                   "t_step=6 s_step=16",
                       "frames": L, "sampler_steps": SAMPLER_STEPS, "windows_per_gpu": n_win_local,
                       "chunk_windows": rt.engine.max_windows, "parallelism": f"time-shard x{world}",
                       "l2": "per-step working set (activations of a chunk + 144 MB packed weights) exceeds the 126 MB "
                             "L2; no explicit flush"},
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "roofline": roofline, "flops_per_step": F_WIN * n_win_local * world,
            "achieved_tflops_whole_step": round(F_WIN * n_win_local / (ms_step * 1e-3) / 1e12, 1),
        }
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ====================================================================================================== reference arm
def cpu_reference(sample_windows: int = 13, steps: int = 1):
    """Times the oracle port of the reference path (fp32 torch on the host cores, all threads) on a bounded sample of
    the same workload: `sample_windows` windows (L = sample_windows + 12 frames), `steps` guided predictor steps, and
    scales to config 2 by the exact work ratio (cost is linear in windows, src/thor/score.py:143-185)."""
    import torch

    from oracle import pipeline_ref, score_ref, unet_ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Ls = sample_windows + 2 * K_ORDER
    sd = unet_ref.init_state_dict(unet_ref.SDA_UNET, seed=0)
    net = unet_ref.RefNet(sd, unet_ref.SDA_UNET)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(Ls, C, H, W, generator=g)
    y = score_ref.coarse_grain(torch.randn(Ls, C, H, W, generator=g), T_STEP, S_STEP)
    std = torch.tensor(STD).reshape(1, C, 1, 1)
    p = pipeline_ref.RefPipeline()
    ts = torch.linspace(1, 0, SAMPLER_STEPS + 1)

    def guided(xx, tt):
        with torch.no_grad():
            eps = score_ref.window_score(net, xx, tt, K_ORDER, batch_size=32)
        return score_ref.guided_score_closed_form(eps, xx, tt, y, std, GAMMA, T_STEP, S_STEP)

    t0 = time.perf_counter()
    for i in range(steps):
        x = p.predictor(guided, x, ts[i], 1 / SAMPLER_STEPS)
    dt = (time.perf_counter() - t0) / steps
    sec_per_window = dt / sample_windows
    L = 2 * K_ORDER + WINDOWS_PER_GPU
    value = L / (SAMPLER_STEPS * sec_per_window * WINDOWS_PER_GPU)
    return {"value": round(value, 6), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} guided predictor step(s) on L={Ls} ({sample_windows} windows), scaled x"
                      f"{WINDOWS_PER_GPU}/{sample_windows} windows to config 2 (cost linear in windows)",
            "sec_per_window_eval": round(sec_per_window, 4), "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per = []
    for _ in range(args.warmup_ref):
        cpu_reference(sample_windows=args.cpu_windows, steps=1)
    steps = max(1, min(args.steps, args.ref_steps))
    for _ in range(steps):
        per.append(cpu_reference(sample_windows=args.cpu_windows, steps=1))
    vals = [p["value"] for p in per]
    value = len(vals) / sum(1.0 / v for v in vals)  # harmonic mean == total work / total time
    cpu = dict(per[-1])
    cpu["value"] = round(value, 6)
    L = 2 * K_ORDER + WINDOWS_PER_GPU
    line = {"impl": "reference", "metric": "guided-sampling frames/sec", "value": round(value, 6), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup_ref,
            "ms_per_step": round(1e3 * L / (SAMPLER_STEPS * value), 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config2: guided PC sampling, L={L} frames, sda_unet.yml ScoreUNet k={K_ORDER}, "
                                   f"{SAMPLER_STEPS} steps, 0 corrections (CPU oracle port of the reference path; each "
                                   f"step = bounded sample of {args.cpu_windows} windows, scaled by the window ratio)"},
            "cpu_baseline": cpu,
            "e2e": {"value": round(value, 6), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=None, help="override L (default 12 + 156 * gpus)")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("C2W_CHUNK", 156)), help="windows per UNet launch")
    ap.add_argument("--e2e-steps", type=int, default=64,
                    help="denoising steps of the end-to-end sample() call (its fixed costs — noise upload, final gather and "
                         "download — are amortised over these steps; the real run has 256)")
    ap.add_argument("--cpu-windows", type=int, default=13)
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--exact-grad", action="store_true",
                    help="guidance through the UNet vector-Jacobian product (condition_on(exact_grad=True)); the shipped "
                         "experiment configs and the default bench line use the closed-form guidance")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: no warm-up floor, no roofline/e2e/cpu legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.profile:
            args.no_e2e = args.no_cpu = True
        else:
            args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
