#!/usr/bin/env python
"""Where a denoising step's time goes, measured in the steady state (no profiler, no events between kernels):

    python -m climate2weather_b200.build --diag
    C2W_LIB=climate2weather_b200/libc2w_b200_diag.so python tools/timeline.py [--steps 6] [--out profiles/rXX_timeline.txt]

The diagnostics build makes every K1 launch stamp {first CTA past its prologue, last CTA done} with %globaltimer.  Over
the timed steps of config 2 (156 windows) this prints: time inside K1, the gaps between consecutive K1 launches split by
what ran in between (nothing = launch / drain / fill overhead; a LayerNorm, attention, gather, guidance kernel), and
the CUDA-event time of the same steps."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    assert os.environ.get("C2W_LIB"), "set C2W_LIB to the diagnostics library"
    from climate2weather_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    net, noise, y, cg = bench.make_problem(bench.L_WEEK)
    st = bench.Stepper(net.to(dev), noise, y, cg, dev, 192, False, shard=False)
    st.load()
    for i in range(4):
        st.step(i)
    cap = 200 * a.steps
    buf = torch.empty(cap, 2, dtype=torch.int64, device=dev)
    buf[:, 0] = torch.iinfo(torch.int64).max
    buf[:, 1] = 0
    eng = st.rt.engine
    _lib.check(lib.c2w_set_timeline(eng.handle, buf.data_ptr(), cap), "c2w_set_timeline")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        st.step(i)
    e1.record()
    torch.cuda.synchronize()
    used = lib.c2w_set_timeline(eng.handle, None, 0)
    t = buf[:used].cpu().numpy().astype("float64")
    per_step = used // a.steps
    inside = (t[:, 1] - t[:, 0]).sum() / 1e6 / a.steps
    gaps = (t[1:, 0] - t[:-1, 1]) / 1e3  # us, between consecutive K1 launches
    step_edges = [(i + 1) * per_step - 1 for i in range(a.steps - 1)]  # gap across a step boundary: K6 + K0 + modulation
    inner = [g for i, g in enumerate(gaps) if i not in step_edges]
    edge = [gaps[i] for i in step_edges]
    wall = (t[-1, 1] - t[0, 0]) / 1e6 / a.steps
    ev = e0.elapsed_time(e1) / a.steps
    import numpy as np
    inner = np.array(inner)
    small = inner[inner < 8.0]
    big = inner[inner >= 8.0]
    lines = [
        f"config 2 (156 windows), {a.steps} steps, diagnostics build, steady state (no events between kernels)",
        f"K1 launches per step: {per_step}",
        f"CUDA-event time per step:                 {ev:8.3f} ms",
        f"first K1 start -> last K1 end per step:   {wall:8.3f} ms",
        f"inside K1 (sum of start..end):            {inside:8.3f} ms  ({100 * inside / ev:.1f} % of the step)",
        f"gaps between consecutive K1 launches with nothing in between ({len(small) // a.steps} per step, < 8 us): "
        f"{small.sum() / 1e3 / a.steps:.3f} ms per step, median {np.median(small):.2f} us, p90 {np.percentile(small, 90):.2f} us",
        f"gaps holding another kernel (LayerNorm / attention, {len(big) // a.steps} per step): {big.sum() / 1e3 / a.steps:.3f} ms per step",
        f"step boundary (guidance + predictor K6, halo, modulation K3, gather K0): {np.mean(edge) / 1e3:.3f} ms per step",
    ]
    text = "\n".join(lines)
    print(text)
    if a.out:
        Path(a.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
