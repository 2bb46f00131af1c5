"""OPT-IN tests of round-2 groundwork (C2W_EXPERIMENTAL=1 pytest -m gpu ...).  Skipped by default so that unfinished
kernels can neither pass silently nor break the suite; nothing here is covered by a parity claim.  State at the end of
round 1: the transpose tests pass on B200, every wgrad case fails with cudaErrorIllegalInstruction.

  c2w_op_wgrad / c2w_op_transpose_bf16 (csrc/wgrad_tcgen05.cuh): weight gradient of a 3x3 stride-1 conv as a split-K
  tcgen05 GEMM, against torch's conv2d weight gradient on the same bf16 operands.
"""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("C2W_EXPERIMENTAL") != "1", reason="opt-in: C2W_EXPERIMENTAL=1")]


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@pytest.mark.parametrize("rows,cols", [(64, 64), (1000, 192), (4096, 128)])
def test_transpose_bf16(rows, cols):
    from climate2weather_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    a = torch.randn(rows, cols, device=dev).to(torch.bfloat16)
    out = torch.empty(cols, rows, device=dev, dtype=torch.bfloat16)
    _lib.check(lib.c2w_op_transpose_bf16(a.data_ptr(), out.data_ptr(), rows, cols, _stream(dev)), "transpose")
    torch.cuda.synchronize()
    assert torch.equal(out, a.t().contiguous())


@pytest.mark.parametrize("n,H,W,cin,cout", [(2, 16, 16, 64, 128), (3, 32, 32, 128, 128), (2, 128, 128, 128, 128),
                                            (5, 8, 8, 256, 384), (2, 64, 64, 128, 64)])
def test_wgrad_vs_torch(n, H, W, cin, cout):
    from climate2weather_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n * 7 + cin)
    x = torch.randn(n, H, W, cin, generator=g).to(dev).to(torch.bfloat16)       # NHWC
    dy = torch.randn(n, H, W, cout, generator=g).to(dev).to(torch.bfloat16)
    pix = n * H * W
    xt = torch.empty(cin, pix, device=dev, dtype=torch.bfloat16)
    dyt = torch.empty(cout, pix, device=dev, dtype=torch.bfloat16)
    st = _stream(dev)
    _lib.check(lib.c2w_op_transpose_bf16(x.data_ptr(), xt.data_ptr(), pix, cin, st), "transpose x")
    _lib.check(lib.c2w_op_transpose_bf16(dy.data_ptr(), dyt.data_ptr(), pix, cout, st), "transpose dy")
    dw = torch.zeros(cout, 9 * cin, device=dev)
    _lib.check(lib.c2w_op_wgrad(xt.data_ptr(), dyt.data_ptr(), n, H, W, cin, cout, dw.data_ptr(), st), "c2w_op_wgrad")
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
    y = F.conv2d(xr, w, padding=1)
    (gw,) = torch.autograd.grad(y, w, dy.float().permute(0, 3, 1, 2))
    want = gw.permute(0, 2, 3, 1).reshape(cout, 9 * cin)  # k = (r*3+s)*cin + ci
    err = (dw - want).abs().max().item() / want.abs().max().item()
    assert err < 1e-3, err  # fp32 accumulation of exact bf16 products, different summation order
