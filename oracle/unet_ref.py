"""Oracle: functional fp32 restatement of the reference ScoreUNet forward (TEST INFRASTRUCTURE ONLY).

Follows model/score.py:14-70 (time embedding, ScoreUNet.forward) and model/nn.py:18-28 (modulated
residual block), :31-85 (attention), :108-218 (layer construction order and names), :220-242
(UNet.forward).  Weights are a plain dict keyed exactly like the reference state_dict
(SURVEY.md §8(b)), so a reference module's `state_dict()` can be fed in directly.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SDA_UNET = dict(  # configs/sda_unet.yml:1-17 + train.py:164-173 (channels = 4 vars * 13 frames)
    channels=52,
    embedding_dim=512,
    hidden_channels=(128, 128, 256, 384, 512),
    hidden_blocks=(3, 3, 3, 3, 3),
    attention_levels=(4,),
    kernel_size=3,
)
NOISE_FEATURES = 32  # model/score.py:53


def param_specs(cfg: dict) -> List[Tuple[str, str, Tuple[int, ...]]]:
    """(state_dict prefix, layer kind, weight shape) in the order the reference constructs them.

    Order matters: it is the order the torch default initialisers consume the global RNG
    (model/score.py:44-57: UNet first, then map_layer0, map_layer1; model/nn.py:165-212: per level
    head, tail, then per block descent, ascent[, descent attention, ascent attention]).
    Names: `tails` and `ascent` are stored reversed (model/nn.py:216,218).
    """
    ch: Sequence[int] = cfg["hidden_channels"]
    blocks: Sequence[int] = cfg["hidden_blocks"]
    attn = set(cfg.get("attention_levels", ()))
    emb = cfg["embedding_dim"]
    cin = cfg["channels"]
    ks = cfg.get("kernel_size", 3)
    nl = len(blocks)
    out: List[Tuple[str, str, Tuple[int, ...]]] = []
    if cfg.get("forcing_dim", 0) > 0:  # model/score.py:49-51: constructed before the UNet
        out.append(("map_forcing", "linear", (emb, cfg["forcing_dim"])))
    for lvl in range(nl):
        rev = nl - 1 - lvl
        if lvl > 0:
            out.append((f"unet.heads.{lvl}.0", "conv", (ch[lvl], ch[lvl - 1], ks, ks)))
            out.append((f"unet.tails.{rev}.2", "conv", (ch[lvl - 1], ch[lvl], ks, ks)))
        else:
            out.append(("unet.heads.0", "conv", (ch[0], cin, ks, ks)))
            out.append((f"unet.tails.{rev}", "conv", (cin, ch[0], ks, ks)))
        stride = 2 if lvl in attn else 1  # attention blocks interleave with residual blocks
        for b in range(blocks[lvl]):
            for side, idx in (("descent", lvl), ("ascent", rev)):
                p = f"unet.{side}.{idx}.{b * stride}"
                out.append((p + ".project.0", "linear", (ch[lvl], emb)))
                out.append((p + ".residue.1", "conv", (ch[lvl], ch[lvl], ks, ks)))
                out.append((p + ".residue.3", "conv", (ch[lvl], ch[lvl], ks, ks)))
            if lvl in attn:
                for side, idx in (("descent", lvl), ("ascent", rev)):
                    p = f"unet.{side}.{idx}.{b * stride + 1}"
                    out.append((p + ".qkv", "conv1d", (3 * ch[lvl], ch[lvl], 1)))
                    out.append((p + ".proj_out", "conv1d", (ch[lvl], ch[lvl], 1)))
    out.append(("map_layer0", "linear", (emb, NOISE_FEATURES)))
    out.append(("map_layer1", "linear", (emb, emb)))
    return out


def init_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init weights identical to `torch.manual_seed(seed); ScoreUNet(**cfg)` in the reference:
    torch's default Conv/Linear initialiser (kaiming_uniform(a=sqrt(5)) weight, U(+-1/sqrt(fan_in)) bias)
    applied in construction order."""
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, _kind, shape in param_specs(cfg):
        w = torch.empty(shape)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        fan_in = w[0].numel()
        bound = 1.0 / math.sqrt(fan_in)
        b = torch.empty(shape[0]).uniform_(-bound, bound)
        sd[name + ".weight"] = w
        sd[name + ".bias"] = b
    return sd


def timestep_embedding(t: Tensor, dim: int = NOISE_FEATURES, max_period: float = 10000.0) -> Tensor:
    """model/score.py:14-34 — [cos(t f_j), sin(t f_j)], f_j = exp(-ln(max_period) j / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.reshape(-1, 1).float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def channel_layernorm(x: Tensor, dim: int = 1, eps: float = 1e-5) -> Tensor:
    """zuko.nn.LayerNorm(dim) restated (parity unpinned, see oracle/__init__.py): standardise over the
    channel axis with the unbiased variance, no affine (model/nn.py:44,154,183)."""
    var, mean = torch.var_mean(x, dim=dim, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def time_modulation(sd: Dict[str, Tensor], t: Tensor, forcing=None) -> Tensor:
    """model/score.py:61-67: emb = silu(W1 silu(W0 e + b0) + b1 [+ Wf forcing + bf])."""
    e = timestep_embedding(t.reshape(-1))
    e = F.silu(F.linear(e, sd["map_layer0.weight"], sd["map_layer0.bias"]))
    e = F.linear(e, sd["map_layer1.weight"], sd["map_layer1.bias"])
    if "map_forcing.weight" in sd:  # model/score.py:65-66 (asserted: forcing is None only without the branch, :60)
        e = e + F.linear(forcing, sd["map_forcing.weight"], sd["map_forcing.bias"])
    else:
        assert forcing is None
    return F.silu(e)


def _res_block(sd, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """model/nn.py:27-28 with residue = LN, conv, SiLU, conv (model/nn.py:151-158)."""
    mod = F.linear(emb, sd[p + ".project.0.weight"], sd[p + ".project.0.bias"])[:, :, None, None]
    h = channel_layernorm(x + mod)
    h = F.conv2d(h, sd[p + ".residue.1.weight"], sd[p + ".residue.1.bias"], padding=1)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".residue.3.weight"], sd[p + ".residue.3.bias"], padding=1)
    return x + h


def _attention(sd, p: str, x: Tensor) -> Tensor:
    """model/nn.py:50-85: single head, tokens = pixels, scale c^-1/4 on q and k, fp32 softmax."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, hh * ww)
    qkv = F.conv1d(channel_layernorm(xf), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    q, k, v = torch.split(qkv, c, dim=1)
    s = 1.0 / math.sqrt(math.sqrt(c))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v)
    a = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + a).reshape(b, c, hh, ww)


def _level(sd, cfg, side: str, lvl: int, x: Tensor, emb: Tensor) -> Tensor:
    nl = len(cfg["hidden_blocks"])
    idx = lvl if side == "descent" else nl - 1 - lvl
    has_attn = lvl in set(cfg.get("attention_levels", ()))
    step = 2 if has_attn else 1
    for b in range(cfg["hidden_blocks"][lvl]):
        x = _res_block(sd, f"unet.{side}.{idx}.{b * step}", x, emb)
        if has_attn:
            x = _attention(sd, f"unet.{side}.{idx}.{b * step + 1}", x)
    return x


def score_unet_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, t: Tensor, forcing=None) -> Tensor:
    """ScoreUNet.forward (model/score.py:59-70) -> UNet.forward (model/nn.py:220-242)."""
    emb = time_modulation(sd, t, forcing)
    nl = len(cfg["hidden_blocks"])
    skips = []
    h = x
    for lvl in range(nl):
        if lvl == 0:
            h = F.conv2d(h, sd["unet.heads.0.weight"], sd["unet.heads.0.bias"], padding=1)
        else:
            h = F.conv2d(h, sd[f"unet.heads.{lvl}.0.weight"], sd[f"unet.heads.{lvl}.0.bias"], stride=2, padding=1)
        h = _level(sd, cfg, "descent", lvl, h, emb)
        skips.append(h)
    skips.pop()
    for lvl in reversed(range(nl)):
        rev = nl - 1 - lvl
        h = _level(sd, cfg, "ascent", lvl, h, emb)
        if lvl > 0:
            u = F.interpolate(channel_layernorm(h), scale_factor=2, mode="nearest")
            h = F.conv2d(u, sd[f"unet.tails.{rev}.2.weight"], sd[f"unet.tails.{rev}.2.bias"], padding=1) + skips.pop()
        else:
            h = F.conv2d(h, sd[f"unet.tails.{rev}.weight"], sd[f"unet.tails.{rev}.bias"], padding=1)
    return h.reshape(x.shape)


class RefNet:
    """Callable (x, t) -> eps with the reference module's call signature, over a weight dict."""

    def __init__(self, sd: Dict[str, Tensor], cfg: dict):
        self.sd, self.cfg = sd, cfg

    def eval(self):
        return self

    def __call__(self, x: Tensor, t: Tensor, forcing=None) -> Tensor:
        return score_unet_forward(self.sd, self.cfg, x, t, forcing)
