"""K10 (weight-gradient GEMM on tcgen05, MN-major operands) and the column-sum kernel, op level, against torch autograd
(pytest -m gpu).  bf16 operands, fp32 accumulation over up to ~2.6 M pixels, deterministic split-K reduction:
tolerance 2^-8 of the gradient's max-abs and 5e-3 relative L2 against fp32 autograd on the SAME bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _wgrad(lib, x_nhwc, dy_nhwc, stride, conv3x3, cin, cout, accumulate=False, into=None, db=None):
    from climate2weather_b200 import _lib
    dev = x_nhwc.device
    n, H, W, cin_pad = x_nhwc.shape
    cout_pad = dy_nhwc.shape[-1]
    taps = 9 if conv3x3 else 1
    scratch = torch.empty(48 * 1024 * 1024, device=dev)  # 192 MB of fp32 partial sums
    dw = into if into is not None else torch.full((cout, cin, taps), 7.0, device=dev)
    _lib.check(lib.c2w_op_wgrad(x_nhwc.data_ptr(), dy_nhwc.data_ptr(), n, H, W, cin_pad, cout_pad, stride, int(conv3x3),
                                scratch.data_ptr(), scratch.numel(), dw.data_ptr(), db.data_ptr() if db is not None else None,
                                cin, cout, int(accumulate), _stream()),
               "c2w_op_wgrad")
    torch.cuda.synchronize()
    return dw


@pytest.mark.parametrize("n,H,W,cin,cout,stride", [
    (2, 16, 16, 64, 64, 1), (3, 32, 32, 128, 128, 1), (2, 128, 128, 128, 128, 1), (2, 128, 128, 52, 128, 1),
    (1, 128, 128, 128, 52, 1), (5, 8, 8, 512, 512, 1), (3, 16, 16, 384, 384, 1), (2, 32, 32, 384, 256, 1),
    (2, 64, 64, 128, 128, 2), (3, 32, 32, 256, 384, 2), (4, 16, 16, 384, 512, 2), (7, 8, 8, 64, 192, 1)])
def test_wgrad_conv3x3_vs_autograd(n, H, W, cin, cout, stride):
    from climate2weather_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n * 100 + cin + cout + stride)
    cin_pad, cout_pad = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    Ho, Wo = H // stride, W // stride
    x = torch.zeros(n, H, W, cin_pad, dtype=torch.bfloat16, device=dev)
    x[..., :cin] = torch.randn(n, H, W, cin, generator=g).to(dev).to(torch.bfloat16)
    dy = torch.zeros(n, Ho, Wo, cout_pad, dtype=torch.bfloat16, device=dev)
    dy[..., :cout] = torch.randn(n, Ho, Wo, cout, generator=g).to(dev).to(torch.bfloat16)
    db = torch.full((cout,), 7.0, device=dev)
    got = _wgrad(lib, x, dy, stride, True, cin, cout, db=db).reshape(cout, cin, 3, 3)
    want_db = dy[..., :cout].float().sum(dim=(0, 1, 2))  # the bias gradient rides along with the GEMM
    assert ((db - want_db).abs().max() / want_db.abs().max()).item() < 1e-5
    w = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
    y = F.conv2d(x[..., :cin].float().permute(0, 3, 1, 2), w, None, stride=stride, padding=1)
    (want,) = torch.autograd.grad(y, w, dy[..., :cout].float().permute(0, 3, 1, 2))
    err = ((got - want).abs().max() / want.abs().max()).item()
    l2 = ((got - want).norm() / want.norm()).item()
    print(f"\nwgrad n={n} {H}x{W} {cin}->{cout} s{stride}: max-abs ratio {err:.2e}, rel-L2 {l2:.2e}")
    assert err < 2 ** -8 and l2 < 5e-3
    # accumulate mode adds to what is there; the reduction is deterministic (bit-identical repeat)
    twice = _wgrad(lib, x, dy, stride, True, cin, cout, accumulate=True, into=got.reshape(cout, cin, 9).clone())
    again = _wgrad(lib, x, dy, stride, True, cin, cout)
    assert torch.equal(again.reshape(cout, cin, 3, 3), got)
    assert torch.allclose(twice.reshape(cout, cin, 3, 3), 2 * got, rtol=1e-6, atol=0)


@pytest.mark.parametrize("rows,cin,cout", [(64, 64, 64), (64 * 9, 512, 1536), (64 * 5, 512, 512), (128, 128, 192)])
def test_wgrad_gemm_vs_torch(rows, cin, cout):
    """1x1 Conv1d of the attention blocks (model/nn.py:45,47): dW = dY^T X."""
    from climate2weather_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(rows + cin)
    x = torch.randn(rows, cin, generator=g).to(dev).to(torch.bfloat16)
    dy = torch.randn(rows, cout, generator=g).to(dev).to(torch.bfloat16)
    db = torch.zeros(cout, device=dev)
    got = _wgrad(lib, x.reshape(1, 1, rows, cin), dy.reshape(1, 1, rows, cout), 1, False, cin, cout, db=db).reshape(cout, cin)
    assert torch.allclose(db, dy.float().sum(0), rtol=1e-5, atol=1e-4)
    want = dy.float().t() @ x.float()
    assert ((got - want).abs().max() / want.abs().max()).item() < 2 ** -8
    assert ((got - want).norm() / want.norm()).item() < 5e-3


@pytest.mark.parametrize("rows,C,rpg", [(1000, 128, 1000), (4096, 512, 1024), (777, 64, 100), (2 * 16384, 128, 16384), (640, 1536, 640), (300, 1000, 64)])
def test_colsum_vs_torch(rows, C, rpg):
    """Bias gradients (one group) and per-sample modulation gradients (one group per image)."""
    from climate2weather_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    x = torch.randn(rows, C, generator=torch.Generator().manual_seed(rows)).to(dev).to(torch.bfloat16)
    groups = -(-rows // rpg)
    out = torch.zeros(groups, C + 8, device=dev)  # out_stride > C: rows of a wider matrix
    _lib.check(lib.c2w_op_colsum(x.data_ptr(), out.data_ptr(), rows, C, rpg, C + 8, 0.5, _stream()), "colsum")
    torch.cuda.synchronize()
    want = torch.stack([x[i * rpg:(i + 1) * rpg].float().sum(0) for i in range(groups)]) * 0.5
    assert torch.allclose(out[:, :C], want, rtol=1e-4, atol=1e-3)
    assert float(out[:, C:].abs().max()) == 0.0
