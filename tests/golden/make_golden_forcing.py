"""Generates tests/golden/forcing.npz by running the REFERENCE's ScoreUNet with a forcing branch (model/score.py:46-67,
forcing_dim = 3) in the build container: forward with one diffusion time and one forcing row per sample, and the
parameter gradients of a squared-error loss under torch autograd.

    python tests/golden/make_golden_forcing.py

No shipped experiment config uses forcing (train.py:167-173 passes forcing_dim = 0); the fixture pins the branch for
the oracle and the CUDA path anyway.  No reference source is copied.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as mg  # noqa: E402


def main():
    ScoreUNet, _, _ = mg.load_reference()
    torch.set_num_threads(8)
    torch.manual_seed(3)
    net = ScoreUNet(channels=20, forcing_dim=3, spatial=2, activation=torch.nn.SiLU, **mg.SMALL).train()
    names = [n for n, _ in net.named_parameters()]
    g = torch.Generator().manual_seed(8)
    x = torch.randn(3, 20, 32, 32, generator=g)
    t = torch.rand(3, 1, 1, 1, generator=g)
    forcing = torch.randn(3, 3, generator=g)
    eps = torch.randn(3, 20, 32, 32, generator=g)
    out = net(x, t, forcing=forcing)
    loss = ((out - eps) ** 2).mean()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    np.savez_compressed(
        mg.OUT / "forcing.npz", names=np.array(names), x=x.numpy(), t=t.numpy(), forcing=forcing.numpy(), eps=eps.numpy(),
        out=out.detach().numpy(), loss=float(loss.item()),
        w_sum=np.array([p.detach().double().sum().item() for p in net.parameters()]),
        grad_sq=np.array([grads[n].double().pow(2).sum().item() for n in names]),
        **{"g::" + n: grads[n].numpy() for n in ("map_forcing.weight", "map_forcing.bias", "map_layer1.weight",
                                                 "unet.descent.0.0.project.0.weight")})
    print("loss", float(loss.item()), "params", len(names))


if __name__ == "__main__":
    main()
