"""The "library bar" of SURVEY.md §8(d) (pytest -m gpu): the reference's own formulation of the ScoreUNet forward —
the oracle's plain torch ops — run ON THE SAME B200 through cuDNN / cuBLAS with bf16 autocast and channels_last
activations, timed beside this package's tcgen05 path on the same weights and inputs.  The hand-written path has to
beat the library path; the measured ratio is printed (and recorded in DESIGN.md).  Both outputs are also held to the
usual tolerance against each other's fp32 ground truth, so the timing compares like with like.
"""
import pytest
import torch

from oracle import unet_ref

pytestmark = pytest.mark.gpu


def _time(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_unet_forward_beats_cudnn_bf16_path():
    import climate2weather_b200 as c2w

    dev = torch.device("cuda:0")
    cfg = unet_ref.SDA_UNET
    torch.manual_seed(0)
    net = c2w.ScoreUNet(52, 512, hidden_channels=[128, 128, 256, 384, 512], hidden_blocks=[3] * 5, attention_levels=[4], activation=torch.nn.SiLU)
    sd_dev = {k: v.detach().to(dev) for k, v in net.state_dict().items()}
    lib_net = unet_ref.RefNet(sd_dev, cfg)  # the oracle's torch-op forward with the weights on the GPU
    net = net.to(dev)
    n = 32
    x = torch.randn(n, 52, 128, 128, generator=torch.Generator().manual_seed(1)).to(dev)
    xl = x.contiguous(memory_format=torch.channels_last)
    t = torch.tensor(0.5, device=dev)
    torch.backends.cudnn.benchmark = True

    def ours():
        with torch.no_grad():
            return net(x, t)

    def library():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return lib_net(xl, t)

    net.DEFAULT_MAX_WINDOWS = n
    y_ours, y_lib = ours(), library().float()
    with torch.no_grad():
        y32 = lib_net(x[:2], t)  # fp32 ground truth on two windows
    scale = y32.abs().max()
    assert ((y_ours[:2] - y32).abs().max() / scale).item() < 3e-2
    assert ((y_lib[:2] - y32).abs().max() / scale).item() < 6e-2
    ms_ours, ms_lib = _time(ours), _time(library)
    flop = 116.0e9 * n
    print(f"\nScoreUNet forward, {n} windows: this package {ms_ours:.2f} ms ({flop / ms_ours / 1e9:.0f} TFLOP/s incl. "
          f"layout conversion), torch/cuDNN bf16 autocast channels_last {ms_lib:.2f} ms "
          f"({flop / ms_lib / 1e9:.0f} TFLOP/s): {ms_lib / ms_ours:.2f}x")
    assert ms_ours < ms_lib
