"""Training step parity (pytest -m gpu; SURVEY.md §8(f) N2, BASELINE config 5): the reference's own step body

    loss = pipeline.loss(net, x).mean(); loss.backward(); optimizer.step(); ema.update()      (training_loop.py:372-390)

on this package's ScoreUNet — forward with per-sample diffusion times, input-gradient convs (K1), weight-gradient GEMMs
(K10), bias / modulation / time-MLP gradients — against

  * tests/golden/train_step.npz: loss, per-parameter gradient checksums and seven full gradient tensors written by the
    REFERENCE's modules under torch autograd (tests/golden/make_golden_train.py), all 64 parameter tensors;
  * the fp32 oracle's autograd gradients, tensor by tensor (small and full architecture).

Tolerance: bf16 tensor-core operands in the forward AND both backward GEMMs against fp32 autograd: relative L2 error
per parameter tensor <= 5e-2 (observed ~1e-2), loss <= 2e-3 relative.
"""
import numpy as np
import pytest
import torch

from oracle import optim_ref, pipeline_ref, unet_ref

pytestmark = pytest.mark.gpu

SMALL = dict(channels=20, embedding_dim=64, hidden_channels=(64, 128), hidden_blocks=(1, 2), attention_levels=(1,),
             kernel_size=3)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _oracle_grads(cfg, sd, x, t, eps):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    net = unet_ref.RefNet(sd, cfg)
    loss = pipeline_ref.RefPipeline().loss(net, x, t=t, eps=eps).mean()
    grads = torch.autograd.grad(loss, list(sd.values()))
    return loss.detach(), dict(zip(sd.keys(), grads))


def _our_step(net, x, t, eps, dev):
    import climate2weather_b200 as c2w
    pipe = c2w.SDAPipeline()
    xt = pipe.mu(t) * x + pipe.sigma(t) * eps  # src/thor/pipelines.py:22-25 with the injected draws
    out = net(xt.to(dev), t.to(dev))
    loss = ((out - eps.to(dev)) ** 2).mean()
    loss.backward()
    return loss.detach()


def test_training_step_vs_reference_golden(golden_dir):
    import climate2weather_b200 as c2w
    from climate2weather_b200 import optim
    dev = torch.device("cuda:0")
    g = np.load(golden_dir / "train_step.npz")
    names = [str(n) for n in g["names"]]
    torch.manual_seed(3)
    net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL)
    assert sorted(n for n, _ in net.named_parameters()) == sorted(names)  # same tensors; the fixture lists them in the
    gi = {n: i for i, n in enumerate(names)}                               # reference's named_parameters() order
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(dev).train()
    x = torch.from_numpy(g["x"])
    torch.manual_seed(7)  # the reference's draws inside loss(): t = rand(B,1,1,1), eps = randn_like(x)
    t = torch.rand(x.shape[0], 1, 1, 1)
    eps = torch.randn_like(x)
    opt = optim.AdamW(net.parameters(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]), betas=(0.9, 0.999))
    ema = optim.StandardEMA(net, rates=[float(g["ema_rate"])])
    opt.fuse_ema(ema)
    opt.zero_grad()
    loss = _our_step(net, x, t, eps, dev)
    print(f"\nloss {loss.item():.6f} vs reference {float(g['loss']):.6f}")
    assert abs(loss.item() - float(g["loss"])) <= 2e-3 * abs(float(g["loss"]))
    want_loss, want = _oracle_grads(SMALL, sd0, x, t, eps)
    worst = ("", 0.0)
    for i, (n, p) in enumerate(net.named_parameters()):
        assert p.grad is not None and p.grad.data_ptr() == opt.grad.data_ptr() + 4 * opt.offsets[i], n  # flat views kept
        e = rel_l2(p.grad, want[n])
        if e > worst[1]:
            worst = (n, e)
        assert e < 5e-2, (n, e)
        # the reference's own checksums (sum of squares is the robust one; the plain sum cancels)
        gs = p.grad.double().pow(2).sum().item()
        assert abs(gs - g["grad_sq"][gi[n]]) <= 0.1 * max(float(g["grad_sq"][gi[n]]), 1e-30), n
    print(f"worst parameter-gradient rel-L2 vs fp32 autograd: {worst[1]:.3e} ({worst[0]}) over {len(names)} tensors")
    for k in [k for k in g.files if k.startswith("g::")]:
        e = rel_l2(dict(net.named_parameters())[k[3:]].grad, torch.from_numpy(g[k]))
        assert e < 5e-2, (k, e)
    # optimizer.step() + ema.update() on those gradients (fused), against the optimiser oracle on the oracle gradients
    ours = [n for n, _ in net.named_parameters()]
    ref = optim_ref.AdamWEMARef([sd0[n] for n in ours], lr=float(g["lr"]), betas=(0.9, 0.999), eps=1e-8,
                                weight_decay=float(g["weight_decay"]), ema_rate=float(g["ema_rate"]))
    ref.step([want[n] for n in ours])
    opt.step()
    ema.update()
    lr = float(g["lr"])
    for i, (n, p) in enumerate(net.named_parameters()):
        d = (p.detach().cpu().double() - ref.p[i]).abs()
        # AdamW's first step moves every weight by ~lr * sign(g): only sign flips of near-zero gradients may differ
        assert d.max().item() <= 2.2 * lr, (n, d.max().item())
        assert d.mean().item() <= 0.2 * lr, (n, d.mean().item())
    # the next forward runs on the UPDATED weights (device-side re-pack, no stale engine)
    with torch.no_grad():
        y1 = net(x.to(dev), t.to(dev))
    fresh = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL)
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    with torch.no_grad():
        y2 = fresh.to(dev)(x.to(dev), t.to(dev))
    assert torch.equal(y1, y2)


def test_direct_gradient_route_matches_autograd_route():
    """optim.AdamW(direct_grads=True): the backward kernels accumulate straight into the optimiser's flat gradient
    buffer (same layout as c2w_param_layout) — identical gradients to the autograd route, accumulation included."""
    import climate2weather_b200 as c2w
    from climate2weather_b200 import optim
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 20, 32, 32, generator=g)
    t = torch.rand(2, 1, 1, 1, generator=g)
    eps = torch.randn(2, 20, 32, 32, generator=g)
    flats = []
    for direct in (False, True):
        torch.manual_seed(1)
        net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL).to(dev)
        opt = optim.AdamW(net.parameters(), lr=1e-3, direct_grads=direct)
        opt.zero_grad()
        _our_step(net, x, t, eps, dev)
        _our_step(net, x, t, eps, dev)  # a second accumulation round
        flats.append(opt.grad.clone())
    assert ((flats[0] - flats[1]).norm() / flats[0].norm()).item() < 1e-5
    assert float(flats[1].abs().max()) > 0


def test_gradient_accumulation_and_input_gradient():
    """Two backward passes accumulate into .grad like autograd does everywhere (training_loop.py:373-378 accumulation
    rounds); the input gradient of the training path agrees with the frozen-weights VJP path."""
    import climate2weather_b200 as c2w
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL).to(dev)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 20, 32, 32, generator=g)
    t = torch.tensor([0.3, 0.3]).reshape(2, 1, 1, 1)
    eps = torch.randn(2, 20, 32, 32, generator=g)
    _our_step(net, x, t, eps, dev)
    g1 = [p.grad.clone() for p in net.parameters()]
    _our_step(net, x, t, eps, dev)
    for a, p in zip(g1, net.parameters()):
        # the modulation / time-MLP sums use fp32 atomics (summation order varies run to run): 1e-3 relative
        assert torch.allclose(p.grad, 2 * a, rtol=1e-3, atol=1e-5 * float(a.abs().max()))
    xg = x.to(dev).requires_grad_(True)
    gout = torch.randn(2, 20, 32, 32, generator=g).to(dev)
    (gin_train,) = torch.autograd.grad(net(xg, torch.tensor(0.3)), xg, gout)
    net.requires_grad_(False)
    (gin_vjp,) = torch.autograd.grad(net(xg, torch.tensor(0.3)), xg, gout)
    assert rel_l2(gin_train, gin_vjp) < 1e-2
    with pytest.raises(RuntimeError):  # a second forward before backward invalidates the first one's stash
        net.requires_grad_(True)
        y1 = net(x.to(dev), t.to(dev))
        net(x.to(dev), t.to(dev))
        y1.sum().backward()


def test_training_step_full_architecture_vs_oracle():
    """configs/sda_unet.yml (72.1 M parameters, 228 tensors), batch 2 at 128 x 128: every parameter gradient against
    the fp32 oracle's autograd."""
    import climate2weather_b200 as c2w
    dev = torch.device("cuda:0")
    cfg = unet_ref.SDA_UNET
    torch.manual_seed(0)
    net = c2w.ScoreUNet(52, 512, hidden_channels=[128, 128, 256, 384, 512], hidden_blocks=[3] * 5, attention_levels=[4],
                        activation=torch.nn.SiLU)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(dev)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 52, 128, 128, generator=g)
    t = torch.tensor([0.15, 0.8]).reshape(2, 1, 1, 1)
    eps = torch.randn(2, 52, 128, 128, generator=g)
    loss = _our_step(net, x, t, eps, dev)
    want_loss, want = _oracle_grads(cfg, sd0, x, t, eps)
    assert abs(loss.item() - want_loss.item()) <= 2e-3 * want_loss.item()
    errs = {n: rel_l2(p.grad, want[n]) for n, p in net.named_parameters()}
    worst = max(errs, key=errs.get)
    tot = rel_l2(torch.cat([p.grad.reshape(-1) for p in net.parameters()]), torch.cat([want[n].reshape(-1) for n in errs]))
    print(f"\nfull architecture: loss {loss.item():.5f} (oracle {want_loss.item():.5f}); parameter gradients rel-L2 overall "
          f"{tot:.3e}, worst tensor {errs[worst]:.3e} ({worst})")
    assert tot < 3e-2 and errs[worst] < 8e-2


def test_forcing_branch_forward_and_gradients(golden_dir):
    """ScoreUNet(forcing_dim=3) (model/score.py:46-67: emb += map_forcing(forcing)): the per-sample forward and the
    training step's gradients — map_forcing's included — against the REFERENCE's own outputs (tests/golden/forcing.npz)
    and the fp32 oracle's autograd."""
    import climate2weather_b200 as c2w
    dev = torch.device("cuda:0")
    g = np.load(golden_dir / "forcing.npz")
    cfg = dict(SMALL, forcing_dim=3)
    torch.manual_seed(3)
    net = c2w.ScoreUNet(activation=torch.nn.SiLU, **cfg)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(dev)
    x, t, f, eps = (torch.from_numpy(g[k]) for k in ("x", "t", "forcing", "eps"))
    with torch.no_grad():
        y = net(x.to(dev), t.to(dev), forcing=f.to(dev))
    e = rel_l2(y, torch.from_numpy(g["out"]))
    print(f"\nforcing forward rel-L2 vs reference: {e:.3e}")
    assert e < 2e-2
    with pytest.raises(ValueError):
        net(x.to(dev), t.to(dev))  # a network with the branch needs forcing, like the reference's assert (model/score.py:60)
    out = net(x.to(dev), t.to(dev), forcing=f.to(dev))
    loss = ((out - eps.to(dev)) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 2e-3 * float(g["loss"])
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    o = unet_ref.score_unet_forward(sd, cfg, x, t, f)
    want = dict(zip(sd, torch.autograd.grad(((o - eps) ** 2).mean(), list(sd.values()))))
    worst = max((rel_l2(p.grad, want[n]), n) for n, p in net.named_parameters())
    print(f"worst parameter-gradient rel-L2 (forcing net): {worst[0]:.3e} ({worst[1]})")
    assert worst[0] < 5e-2
    for k in ("map_forcing.weight", "map_forcing.bias"):
        assert rel_l2(dict(net.named_parameters())[k].grad, torch.from_numpy(g["g::" + k])) < 3e-2, k
