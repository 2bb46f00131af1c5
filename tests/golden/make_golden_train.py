"""Generates tests/golden/train_step.npz by running the REFERENCE's own training-step arithmetic
in the build container: `pipeline.loss(net, x).mean().mul(loss_scaling)` -> backward -> torch.optim.AdamW.step()
(train.py:176-181) -> StandardEMA.update() (training_loop.py:372-390, src/thor/pipelines.py:27-35,
src/thor/ema.py:24-27), on the small architecture of the other fixtures.

    python tests/golden/make_golden_train.py

It is the provenance of the parameter-gradient parity that the wgrad kernels (next round) will be held to; today it
pins the oracle's autograd gradients and the optimiser oracle.  No reference source is copied.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as mg  # noqa: E402

FULL_GRADS = ("unet.heads.0.weight", "unet.heads.1.0.weight", "unet.tails.1.weight", "unet.descent.0.0.project.0.weight",
              "unet.descent.0.0.residue.1.weight", "unet.ascent.1.0.residue.3.weight", "map_layer0.weight")


def main():
    ScoreUNet, score, pipelines = mg.load_reference()
    ema_mod = mg._load("ref_thor_ema", mg.REF / "src/thor/ema.py")
    torch.set_num_threads(8)
    k, C, H, W = 2, 4, 32, 32
    w = 2 * k + 1
    torch.manual_seed(3)
    net = ScoreUNet(channels=C * w, spatial=2, activation=torch.nn.SiLU, **mg.SMALL).train()
    names = [n for n, _ in net.named_parameters()]
    pipe = pipelines.SDAPipeline()
    ema = ema_mod.StandardEMA(net, rates=[0.99])
    opt = torch.optim.AdamW(net.parameters(), lr=2e-4, weight_decay=1e-3, betas=[0.9, 0.999])
    torch.manual_seed(6)
    xw = torch.randn(3, C * w, H, W)
    torch.manual_seed(7)  # loss() draws t = rand(B,1,1,1) then eps = randn_like(x): replayed by the tests
    opt.zero_grad()
    loss = pipe.loss(net=net, x=xw).mean().mul(1.0)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    opt.step()
    ema.update(cur_ndata=3, batch_size=3)
    out = dict(
        names=np.array(names), x=xw.numpy(), loss=float(loss.item()),
        grad_sum=np.array([grads[n].double().sum().item() for n in names]),
        grad_abs=np.array([grads[n].double().abs().sum().item() for n in names]),
        grad_sq=np.array([grads[n].double().pow(2).sum().item() for n in names]),
        p_sum=np.array([p.detach().double().sum().item() for p in net.parameters()]),
        p_abs=np.array([p.detach().double().abs().sum().item() for p in net.parameters()]),
        ema_sum=np.array([p.detach().double().sum().item() for p in ema.emas[0].parameters()]),
        lr=2e-4, weight_decay=1e-3, ema_rate=0.99)
    for n in FULL_GRADS:
        if n in grads:
            out["g::" + n] = grads[n].numpy()
    assert all(n in grads for n in FULL_GRADS), [n for n in FULL_GRADS if n not in grads]
    np.savez_compressed(mg.OUT / "train_step.npz", **out)
    print("loss", out["loss"], "params", len(names), "full grads", [k for k in out if k.startswith("g::")])


if __name__ == "__main__":
    main()
