"""Does replaying the UNet forward from a CUDA graph beat stream launches?  (inter-kernel gap probe; B200 via gpurun)
Captures one c2w_unet_forward call (113 kernel launches at 156 windows) with torch.cuda.CUDAGraph and times replay
against the direct call."""
import sys
import pathlib
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import climate2weather_b200 as c2w  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = c2w.ScoreUNet(52, 512, hidden_channels=[128, 128, 256, 384, 512], hidden_blocks=[3] * 5, attention_levels=[4], activation=torch.nn.SiLU).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 156
eng = net.engine(4, 13, 128, 128, dev, max_windows=n)
x = torch.randn(n, 52, 128, 128, device=dev)


def direct():
    return eng.unet_forward(x, 0.5)


def timeit(fn, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ref = direct()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    direct()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = direct()
g.replay()
torch.cuda.synchronize()
print("graph output equals direct:", torch.equal(out, ref))
for _ in range(2):
    print(f"direct {timeit(direct):.3f} ms   graph replay {timeit(g.replay):.3f} ms")
