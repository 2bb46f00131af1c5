"""N2 groundwork: the reference's own training-step arithmetic — DSM loss -> backward -> AdamW ->
EMA (training_loop.py:372-390) — recorded by tests/golden/make_golden_train.py from the reference modules, against

  * the oracle's autograd parameter gradients (oracle/unet_ref.py + oracle/pipeline_ref.py): the ground truth the
    wgrad kernels of the next round will be tested against, pinned here for all 64 parameter tensors;
  * the optimiser oracle (oracle/optim_ref.py) applied to those gradients: parameters and EMA after one step.
"""
import numpy as np
import pytest
import torch

from oracle import optim_ref, pipeline_ref, unet_ref

SMALL = dict(channels=20, embedding_dim=64, hidden_channels=(64, 128), hidden_blocks=(1, 2), attention_levels=(1,),
             kernel_size=3)


@pytest.fixture(scope="module")
def step(golden_dir):
    g = np.load(golden_dir / "train_step.npz")
    names = [str(n) for n in g["names"]]
    sd = unet_ref.init_state_dict(SMALL, seed=3)
    for n in names:
        sd[n].requires_grad_(True)
    net = unet_ref.RefNet(sd, SMALL)
    x = torch.from_numpy(g["x"])
    torch.manual_seed(7)  # the reference's draws inside loss(): t = rand(B,1,1,1), eps = randn_like(x)
    loss = pipeline_ref.RefPipeline().loss(net, x).mean()
    grads = torch.autograd.grad(loss, [sd[n] for n in names])
    return g, names, sd, loss.detach(), dict(zip(names, grads))


def test_oracle_loss_and_parameter_gradients_match_reference(step):
    g, names, sd, loss, grads = step
    assert abs(loss.item() - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    for i, n in enumerate(names):
        gr = grads[n].double()
        scale = max(float(g["grad_abs"][i]), 1e-12)
        assert abs(gr.sum().item() - g["grad_sum"][i]) <= 2e-4 * scale, n
        assert abs(gr.abs().sum().item() - g["grad_abs"][i]) <= 2e-4 * scale, n
        assert abs(gr.pow(2).sum().item() - g["grad_sq"][i]) <= 5e-4 * max(float(g["grad_sq"][i]), 1e-20), n
    full = [k for k in g.files if k.startswith("g::")]
    assert len(full) == 7
    for k in full:
        want = torch.from_numpy(g[k])
        got = grads[k[3:]]
        assert got.shape == want.shape
        assert (got - want).abs().max().item() <= 3e-4 * want.abs().max().item(), k


def test_optimizer_oracle_reproduces_reference_step(step):
    g, names, sd, _, grads = step
    ref = optim_ref.AdamWEMARef([sd[n].detach() for n in names], lr=float(g["lr"]), betas=(0.9, 0.999), eps=1e-8,
                                weight_decay=float(g["weight_decay"]), ema_rate=float(g["ema_rate"]))
    ref.step([grads[n].detach() for n in names])
    for i, n in enumerate(names):
        scale = max(float(g["p_abs"][i]), 1e-12)
        # AdamW's first step moves every weight by ~lr regardless of the gradient's size, so gradient noise of 1e-4
        # relative barely matters; sign flips of near-zero gradients are what the tolerance allows for
        assert abs(ref.p[i].sum().item() - g["p_sum"][i]) <= 1e-5 * scale + 2e-4 * float(g["lr"]) * ref.p[i].numel(), n
        assert abs(ref.ema[i].sum().item() - g["ema_sum"][i]) <= 1e-5 * scale + 2e-6 * float(g["lr"]) * ref.p[i].numel(), n
