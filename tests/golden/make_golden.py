"""Generates tests/golden/*.npz by running the REFERENCE's own code (schmidtjonathan/Climate2Weather,
mounted read-only at /root/reference) in the build container.  The reference cannot travel to the GPU
box, so its outputs are committed here as small fixtures; this script is the provenance.

    python tests/golden/make_golden.py            # rewrites the fixtures

Shims: `zuko.nn.LayerNorm` is not installed -> stand-in with the published zuko 1.0.x
semantics; `thor/score.py` and `thor/pipelines.py` are loaded by path because `import thor` pulls in
lightning.  No reference source is copied: the modules are imported where they lie.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _install_zuko_standin():
    zuko = types.ModuleType("zuko")
    znn = types.ModuleType("zuko.nn")

    class LayerNorm(torch.nn.Module):
        def __init__(self, dim=-1, eps: float = 1e-5):
            super().__init__()
            self.dim = tuple(dim) if isinstance(dim, (tuple, list)) else (dim,)
            self.eps = eps

        def forward(self, x):
            var, mean = torch.var_mean(x, dim=self.dim, keepdim=True)
            return (x - mean) / (var + self.eps).sqrt()

    LayerNorm.__module__ = "zuko.nn"  # picklable under the real package's name (snapshot fixture)
    LayerNorm.__qualname__ = "LayerNorm"
    znn.LayerNorm = LayerNorm
    zuko.nn = znn
    sys.modules["zuko"] = zuko
    sys.modules["zuko.nn"] = znn


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    _install_zuko_standin()
    sys.path.insert(0, str(REF))
    from model.score import ScoreUNet  # noqa: E402

    score = _load("ref_thor_score", REF / "src/thor/score.py")
    pipelines = _load("ref_thor_pipelines", REF / "src/thor/pipelines.py")
    return ScoreUNet, score, pipelines


FULL = dict(embedding_dim=512, hidden_channels=[128, 128, 256, 384, 512], hidden_blocks=[3, 3, 3, 3, 3],
            attention_levels=[4], kernel_size=3, padding_mode="zeros")
SMALL = dict(embedding_dim=64, hidden_channels=[64, 128], hidden_blocks=[1, 2], attention_levels=[1],
             kernel_size=3, padding_mode="zeros")
STD = [0.1692666615037876, 0.0425178630338289, 0.3268027589410125, 0.3268027589410125]
GAMMA = 0.0007196856730011522


def weight_checksums(net):
    names, sums, asums = [], [], []
    for k, v in net.state_dict().items():
        names.append(k)
        sums.append(v.double().sum().item())
        asums.append(v.double().abs().sum().item())
    return np.array(names), np.array(sums), np.array(asums)


def make_snapshot(ScoreUNet):
    """A network snapshot exactly as training_loop.py:250-266 pickles it (EasyDict with the fp16 EMA module, the
    pipeline object and dataset_kwargs), at a tiny architecture: the fixture for the compat / unpickling test."""
    import pickle

    util = types.ModuleType("util")  # util.py imports lightning; EasyDict restated under the reference's name

    class EasyDict(dict):
        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name)

        def __setattr__(self, name, value):
            self[name] = value

    EasyDict.__module__ = "util"
    EasyDict.__qualname__ = "EasyDict"
    util.EasyDict = EasyDict
    sys.modules["util"] = util
    thor = types.ModuleType("thor")
    sys.modules["thor"] = thor
    spec = importlib.util.spec_from_file_location("thor.pipelines", REF / "src/thor/pipelines.py")
    tp = importlib.util.module_from_spec(spec)
    sys.modules["thor.pipelines"] = tp
    spec.loader.exec_module(tp)
    tiny = dict(embedding_dim=64, hidden_channels=[64, 64], hidden_blocks=[1, 1], attention_levels=[1], kernel_size=3,
                padding_mode="zeros")
    torch.manual_seed(11)
    net = ScoreUNet(channels=12, spatial=2, activation=torch.nn.SiLU, **tiny)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    snap = EasyDict(dataset_kwargs=EasyDict(train=EasyDict(window=3)), pipeline=tp.SDAPipeline())
    snap.ema = net.cpu().eval().requires_grad_(False).to(torch.float16)
    with open(OUT / "snapshot_tiny.pkl", "wb") as f:
        pickle.dump(snap, f)
    x = torch.randn(2, 12, 16, 16, generator=torch.Generator().manual_seed(12))
    net32 = ScoreUNet(channels=12, spatial=2, activation=torch.nn.SiLU, **tiny)
    net32.load_state_dict({k: v.half().float() for k, v in sd.items()})  # what the fp16 snapshot holds
    with torch.no_grad():
        y = net32.eval()(x, torch.tensor(0.3))
    names = np.array(list(sd.keys()))
    np.savez_compressed(OUT / "snapshot_tiny_expect.npz", names=names,
                        sums=np.array([sd[k].half().double().sum().item() for k in sd]), x=x.numpy(), t=0.3, y=y.numpy())
    print("wrote snapshot_tiny.pkl", (OUT / "snapshot_tiny.pkl").stat().st_size, "bytes")


def main():
    ScoreUNet, score, pipelines = load_reference()
    make_snapshot(ScoreUNet)
    torch.set_num_threads(8)

    # ---------------------------------------------------------------- 1. full architecture (sda_unet.yml)
    torch.manual_seed(0)
    net = ScoreUNet(channels=52, spatial=2, activation=torch.nn.SiLU, **FULL).eval()
    names, sums, asums = weight_checksums(net)
    shapes = np.array([",".join(map(str, v.shape)) for v in net.state_dict().values()])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 52, 128, 128, generator=g)
    t = torch.tensor(0.7)
    with torch.no_grad():
        y = net(x, t)
    np.savez_compressed(OUT / "full_arch.npz", names=names, shapes=shapes, sums=sums, asums=asums,
                        out_slice=y[0, :, ::8, ::8].numpy(), out_mean=y.double().mean().item(),
                        out_std=y.double().std().item(), t=0.7, x_seed=1)
    del net

    # ---------------------------------------------------------------- 2. small architecture, whole path
    k, C, H, W, L = 2, 4, 32, 32, 9
    w = 2 * k + 1
    torch.manual_seed(3)
    net = ScoreUNet(channels=C * w, spatial=2, activation=torch.nn.SiLU, **SMALL).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    names, sums, asums = weight_checksums(net)
    pipe = pipelines.SDAPipeline()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(L, C, H, W, generator=g)
    truth = torch.randn(L, C, H, W, generator=g)
    t_step, s_step = 3, 8
    pool = torch.nn.AvgPool2d(s_step, stride=s_step, padding=0)

    def A(z):
        return pool(z[..., ::t_step, :, :, :])

    yobs = A(truth)
    std = torch.tensor(STD).reshape(1, C, 1, 1)
    t = torch.tensor(0.6)
    out = {}
    with torch.no_grad():
        sf = score.DefaultScoreFunction(net, markov_order=k, noise_process=pipe)
        out["eps_default"] = sf(x, t).numpy()
        for bs in (2, 3, 5):
            bf = score.BatchedScoreFunction(net, markov_order=k, noise_process=pipe, batch_size=bs,
                                            device=torch.device("cpu"))
            out[f"eps_batched_{bs}"] = bf(x, t).numpy()
    for exact in (False, True):
        sf = score.DefaultScoreFunction(net, markov_order=k, noise_process=pipe)
        sf.condition_on(A=A, y=yobs, std=std, gamma=GAMMA, exact_grad=exact)
        out[f"eps_guided_{'exact' if exact else 'approx'}"] = sf(x, t).detach().numpy()
    # sampler: steps=3, corrections=1, guided (approx); corrector noise from the global CPU generator
    sf = score.BatchedScoreFunction(net, markov_order=k, noise_process=pipe, batch_size=4, device=torch.device("cpu"))
    sf.condition_on(A=A, y=yobs, std=std, gamma=GAMMA, exact_grad=False)
    torch.manual_seed(5)
    out["sample_c1"] = pipe.sample(sf, x, steps=3, corrections=1, tau=0.5, show_progressbar=False).numpy()
    torch.manual_seed(5)
    out["sample_c0"] = pipe.sample(sf, x, steps=4, corrections=0, tau=0.5, show_progressbar=False).numpy()
    # unguided sampler as in training_loop.py:296-309 (one window)
    sf1 = score.DefaultScoreFunction(net, markov_order=k, noise_process=pipe)
    out["sample_one_window"] = pipe.sample(sf1, x[:w], steps=3, show_progressbar=False).numpy()
    # DSM loss with injected t / eps (pipelines.py:27-35 draws them; we replay the same draws)
    torch.manual_seed(6)
    xw = torch.randn(2, C * w, H, W)
    torch.manual_seed(7)
    with torch.no_grad():
        out["loss"] = pipe.loss(net, xw).numpy()
    out["loss_x"] = xw.numpy()
    ts = torch.linspace(0, 1, 11)
    np.savez_compressed(OUT / "small_path.npz", names=names, sums=sums, asums=asums, x=x.numpy(), yobs=yobs.numpy(),
                        t=0.6, mu=pipe.mu(ts).numpy(), sigma=pipe.sigma(ts).numpy(), ts=ts.numpy(), **out)

    # ---------------------------------------------------------------- 3. index maps (bit-exact integer work)
    class Ident(torch.nn.Module):
        def forward(self, x, t):
            return x

    idx = {}
    for (L_, k_, C_) in [(13, 6, 4), (14, 6, 4), (26, 6, 4), (40, 6, 4), (5, 2, 4), (9, 2, 3), (7, 1, 1)]:
        code = (torch.arange(L_)[:, None, None, None] * 1000 + torch.arange(C_)[None, :, None, None] * 10
                + torch.arange(2)[None, None, :, None] * 2 + torch.arange(2)[None, None, None, :]).float()
        sf = score.DefaultScoreFunction(Ident(), markov_order=k_, noise_process=pipe)
        u = sf.unfold(code)
        idx[f"unfold_{L_}_{k_}_{C_}"] = u.numpy().astype(np.int64)
        nw = L_ - 2 * k_
        wcode = (torch.arange(nw)[:, None, None, None] * 1000 + torch.arange((2 * k_ + 1) * C_)[None, :, None, None]
                 ).float().expand(nw, (2 * k_ + 1) * C_, 1, 1)
        idx[f"fold_{L_}_{k_}_{C_}"] = sf.fold(wcode).numpy().astype(np.int64)
        for bs in (1, 2, 3, 16):
            bf = score.BatchedScoreFunction(Ident(), markov_order=k_, noise_process=pipe, batch_size=bs,
                                            device=torch.device("cpu"))
            idx[f"batched_{L_}_{k_}_{C_}_{bs}"] = bf.score_fn(code, torch.tensor(0.5)).numpy().astype(np.int64)
    np.savez_compressed(OUT / "index_maps.npz", **idx)
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
