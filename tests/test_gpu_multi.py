"""Multi-GPU parity (pytest -m gpu, needs >= 2 devices; skipped on a 1-GPU box): time-sharded guided sampling over
NCCL must reproduce the single-GPU trajectory.  Frame i's score depends only on frames i-k..i+k (src/thor/score.py:68-93)
and every kernel is deterministic per window, so the sharded run is held to BIT-EXACT equality with the unsharded one
(predictor steps; the corrector's global mean(eps^2) is summed in a different order across ranks -> 1e-6).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

SMALL = dict(channels=20, embedding_dim=64, hidden_channels=(64, 128), hidden_blocks=(1, 2), attention_levels=(1,),
             kernel_size=3)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(L):
    import climate2weather_b200 as c2w

    torch.manual_seed(3)
    net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL)
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(L, 4, 32, 32, generator=g)
    y = c2w.CoarseGrain(3, 8)(torch.randn(L, 4, 32, 32, generator=g))
    return net, noise, y


def _sample(net, noise, y, dev, corrections, shard, exact=False):
    import climate2weather_b200 as c2w

    pipe = c2w.SDAPipeline()
    sf = c2w.BatchedScoreFunction(net.to(dev), markov_order=2, noise_process=pipe, batch_size=5, device=dev)
    sf.condition_on(A=c2w.CoarseGrain(3, 8), y=y, std=0.1, gamma=1e-3, exact_grad=exact)
    if shard:
        sf.enable_time_sharding()
    return pipe.sample(sf, noise, steps=4, corrections=corrections, tau=0.5, show_progressbar=False, seed=77)


def _worker(rank, world, port, L, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        net, noise, y = _problem(L)
        for corr in (0, 1):
            out = _sample(net, noise, y, dev, corr, shard=True)
            if rank == 0:
                torch.save(out.cpu(), os.path.join(out_dir, f"sharded_c{corr}.pt"))
        out = _sample(net, noise, y, dev, 0, shard=True, exact=True)
        if rank == 0:
            torch.save(out.cpu(), os.path.join(out_dir, "sharded_exact.pt"))
        # the same run over torch.distributed send/recv instead of the peer-memory mailboxes (c2w_halo_exchange)
        os.environ["C2W_HALO"] = "nccl"
        out = _sample(net, noise, y, dev, 0, shard=True)
        if rank == 0:
            torch.save(out.cpu(), os.path.join(out_dir, "sharded_c0_nccl.pt"))
        os.environ["C2W_HALO"] = "p2p"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_time_sharded_sampling_matches_single_gpu(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    L = 23
    net, noise, y = _problem(L)
    dev = torch.device("cuda:0")
    ref = {c: _sample(net, noise, y, dev, c, shard=False).cpu() for c in (0, 1)}
    mp.spawn(_worker, args=(world, _free_port(), L, str(tmp_path)), nprocs=world, join=True)
    got0 = torch.load(tmp_path / "sharded_c0.pt")
    got1 = torch.load(tmp_path / "sharded_c1.pt")
    assert torch.isfinite(got0).all() and torch.isfinite(got1).all()
    assert torch.equal(got0, ref[0]), float((got0 - ref[0]).abs().max())
    assert torch.equal(torch.load(tmp_path / "sharded_c0_nccl.pt"), ref[0])  # both halo transports: bit-exact
    rel = ((got1 - ref[1]).abs().max() / ref[1].abs().max()).item()
    assert rel < 1e-5, rel
    # exact_grad: the UNet VJP reaches k frames into the neighbours' shards (reverse halo exchange, send-and-add);
    # the fp32 accumulation order over windows differs between the sharded and the unsharded run -> 1e-4
    want = _sample(net, noise, y, dev, 0, shard=False, exact=True).cpu()
    gote = torch.load(tmp_path / "sharded_exact.pt")
    rel = ((gote - want).abs().max() / want.abs().max()).item()
    assert rel < 1e-4, rel
    assert not torch.equal(want, ref[0])


# ------------------------------------------------------------------------------------------------ data-parallel training
def _train_data(rank_or_all):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 20, 32, 32, generator=g)
    t = torch.rand(4, 1, 1, 1, generator=g)
    eps = torch.randn(4, 20, 32, 32, generator=g)
    if rank_or_all is None:
        return x, t, eps
    sl = slice(2 * rank_or_all, 2 * rank_or_all + 2)
    return x[sl], t[sl], eps[sl]


def _train_one_step(net, x, t, eps, dev, group):
    import climate2weather_b200 as c2w
    from climate2weather_b200 import optim

    pipe = c2w.SDAPipeline()
    opt = optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-3, data_parallel_group=group)
    if group != "none":
        opt.broadcast_parameters(0)
    opt.zero_grad()
    xt = pipe.mu(t) * x + pipe.sigma(t) * eps
    loss = ((net(xt.to(dev), t.to(dev)) - eps.to(dev)) ** 2).mean()
    loss.backward()
    opt.step()
    return [p.detach().cpu().clone() for p in net.parameters()], opt.grad.detach().cpu().clone()


def _train_worker(rank, world, port, out_dir):
    import climate2weather_b200 as c2w

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.manual_seed(3 + rank)  # different initial weights per rank: broadcast_parameters must align them
        net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL).to(dev)
        params, grad = _train_one_step(net, *_train_data(rank), dev, None)
        torch.save((params, grad), os.path.join(out_dir, f"train_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_matches_single_process(tmp_path):
    """training_loop.py:116,375-378 (DDP): two ranks with half the batch each and ONE in-place all-reduce (average) of
    the flat gradient buffer inside optimizer.step() must take the same step as one process on the whole batch (the
    loss is a mean, so the average of the two half-batch gradients is the full-batch gradient).  Parameters after the
    step agree to AdamW's sign-flip bound; the averaged gradients to 2e-2 relative L2 (bf16 kernels, different batch
    split); both ranks end with identical parameters bit for bit."""
    import climate2weather_b200 as c2w

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_train_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    p0, g0 = torch.load(tmp_path / "train_rank0.pt")
    p1, g1 = torch.load(tmp_path / "train_rank1.pt")
    assert torch.equal(g0, g1)
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)
    dev = torch.device("cuda:0")
    torch.manual_seed(3)  # rank 0's initial weights
    net = c2w.ScoreUNet(activation=torch.nn.SiLU, **SMALL).to(dev)
    want, gw = _train_one_step(net, *_train_data(None), dev, "none")
    assert ((g0 - gw).norm() / gw.norm()).item() < 2e-2
    for a, b in zip(p0, want):
        assert (a - b).abs().max().item() <= 2.2e-3
