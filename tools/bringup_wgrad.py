"""Timing sweep of K10 (weight-gradient GEMM, csrc/wgrad_tcgen05.cuh) at the training step's layer shapes (batch 128),
CUDA events around 10 back-to-back launches each (GEMM + slab reduction):

    python tools/bringup_wgrad.py            # table of ms / TFLOP/s per shape
    ncu --set full -k regex:wgrad_tcgen05 -c 4 ... python tools/bringup_wgrad.py --ncu
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from climate2weather_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
NO_BIAS = "--no-bias" in sys.argv  # weight gradient only (no bias column sums)


def run(name, n, H, W, cin, cout, stride=1, iters=10, warm=2):
    cin_pad, cout_pad = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    x = torch.randn(n, H, W, cin_pad, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, H // stride, W // stride, cout_pad, device=dev).to(torch.bfloat16)
    scratch = torch.empty(12 * 1024 * 1024 + 65536, device=dev)
    dw = torch.zeros(cout, cin, 9, device=dev)
    db = torch.zeros(cout, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        _lib.check(lib.c2w_op_wgrad(x.data_ptr(), dy.data_ptr(), n, H, W, cin_pad, cout_pad, stride, 1, scratch.data_ptr(),
                                    scratch.numel(), dw.data_ptr(), 0 if NO_BIAS else db.data_ptr(), cin, cout, 0, st), "wgrad")

    for _ in range(warm):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * (H // stride) * (W // stride) * cout_pad * 9 * cin_pad
    print(json.dumps({"shape": name, "n": n, "H": H, "W": W, "cin": cin, "cout": cout, "stride": stride, "ms": round(ms, 4),
                      "tflops": round(flops / ms / 1e9, 1)}), flush=True)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    if "--ncu" in sys.argv:
        run("G2 128->128 @128^2, batch 32", 32, 128, 128, 128, 128, iters=1, warm=1)
        run("G8 256->256 @32^2, batch 128", 128, 32, 32, 256, 256, iters=1, warm=1)
        return
    B = 128
    run("G2  128->128 @128^2", B, 128, 128, 128, 128)
    run("G1   52->128 @128^2", B, 128, 128, 52, 128)
    run("G3  128->52  @128^2", B, 128, 128, 128, 52)
    run("G4  128->128 s2 ->64^2", B, 128, 128, 128, 128, stride=2)
    run("G5  128->128 @64^2", B, 64, 64, 128, 128)
    run("G8  256->256 @32^2", B, 32, 32, 256, 256)
    run("G11 384->384 @16^2", B, 16, 16, 384, 384)
    run("G14 512->512 @8^2", B, 8, 8, 512, 512)


if __name__ == "__main__":
    main()
