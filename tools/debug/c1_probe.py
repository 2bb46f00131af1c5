import sys; sys.path.insert(0, ".")
import numpy as np, torch
import climate2weather_b200 as c2w
from climate2weather_b200.score import _mu_sigma
SMALL = dict(channels=20, embedding_dim=64, hidden_channels=(64, 128), hidden_blocks=(1, 2), attention_levels=(1,), kernel_size=3)
STD = [0.1692666615037876, 0.0425178630338289, 0.3268027589410125, 0.3268027589410125]; GAMMA = 0.0007196856730011522
dev = torch.device("cuda:0")
g = np.load("tests/golden/small_path.npz")
torch.manual_seed(3)
net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL).to(dev)
pipe = c2w.SDAPipeline(); pipe.rng = "reference"
x = torch.from_numpy(g["x"]); y = torch.from_numpy(g["yobs"])
sf = c2w.BatchedScoreFunction(net, markov_order=2, noise_process=pipe, batch_size=4, device=dev)
sf.condition_on(A=c2w.CoarseGrain(3, 8), y=y, std=torch.tensor(STD).reshape(1, 4, 1, 1), gamma=GAMMA, exact_grad=False)
ref = torch.from_numpy(g["sample_c1"])
rl2 = lambda a, b: ((a - b).norm() / b.norm()).item()
s0 = pipe.sample(sf, x, steps=3, corrections=0, tau=0.5, show_progressbar=False)
print("ours 3 predictor steps vs golden c1:", rl2(s0, ref))
torch.manual_seed(5)
s1 = pipe.sample(sf, x, steps=3, corrections=1, tau=0.5, show_progressbar=False)
print("ours c1 vs golden c1:", rl2(s1, ref), " ours c1 vs ours c0:", rl2(s1, s0))
# manual loop with prints
rt = sf.runtime(x); rt.load(x)
ts = torch.linspace(1, 0, 4); dt = 1 / 3
for i, t in enumerate(ts[:-1]):
    mu, sg = _mu_sigma(pipe, t); mun, sgn = _mu_sigma(pipe, t - dt)
    rt.score(float(t)); rt.predictor(mu, sg, mun, sgn)
    xp = rt.owned(rt.x).cpu()
    rt.score(float(t - dt)); rt.guided_eps(mun, sgn)
    eg = rt.owned(rt.eps_g).cpu()
    rt.lib.c2w_reduce_partials(rt.partials.data_ptr(), rt._n_part, rt.sumsq.data_ptr(), rt.stream)
    torch.cuda.synchronize()
    print(f"step {i}: |x| {xp.norm():.4e} sum eps_g^2 torch {eg.double().square().sum():.6e} kernel {rt.sumsq.item():.6e} "
          f"n_part {rt._n_part} partials.numel {rt.partials.numel()} delta {0.5 / (rt.sumsq.item() / eg.numel()):.3e}")
    z = torch.zeros_like(rt.x)
    rt.corrector(0.5, sgn, z, 0, i)
    xc = rt.owned(rt.x).cpu()
    print(f"        corrector moved x by rel {rl2(xc, xp):.3e}")
