"""N4: the step either side of the sampling path — the reference's per-variable normalisation
(`data/pipeline.py:183-247`) fused on the device with the layout change of `ds_to_sorted_np` / `np_to_ds`
(`data/pipeline.py:250-272`) and the trajectory packing, so the observation / ground truth go from per-variable arrays
to the device trajectory layout [L, H, W, C] in one pass, and a finished sample comes back un-normalised per variable in
one pass.  Function names and the `mode` strings are the reference's; datasets are plain `{variable: array[L, H, W]}`
mappings (what `ds[v].values` holds) and quantiles `{q: {variable: scalar | array[H, W]}}` (what
`quantile_ds.sel(quantile=q)[v]` holds) — xarray itself stays with the caller.

There is no CPU path: the arithmetic runs in `c2w_normalize_pack` / `c2w_unpack_unnormalize`.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Sequence, Tuple

import numpy as np
import torch

from . import _lib

#: mode -> (quantile subtracted, (quantiles whose difference divides))      data/pipeline.py:189-213
NORM_MODES = {
    "minmax": (0.0, (0.0, 1.0)),
    "robust": (0.5, (0.25, 0.75)),
    "robust95": (0.5, (0.05, 0.95)),
    "quant95": (0.05, (0.05, 0.95)),
    "quant99": (0.01, (0.01, 0.99)),
}


def _quantile(quantiles: Mapping, q: float):
    for k in quantiles:
        if abs(float(k) - q) < 1e-9:
            return quantiles[k]
    raise KeyError(f"quantile {q} not in the quantile set {sorted(float(k) for k in quantiles)}")


def coefficients(quantiles: Mapping, data_vars: Sequence[str], mode: str) -> Tuple[np.ndarray, np.ndarray]:
    """(shift, scale) as float32 [C] (scalar quantiles) or [C, H, W] (per-grid-point quantiles), variables sorted
    like the reference sorts them (data/pipeline.py:255)."""
    if mode not in NORM_MODES:
        raise ValueError(f"Invalid mode: {mode}")  # data/pipeline.py:214-215
    q_shift, (q_lo, q_hi) = NORM_MODES[mode]
    names = list(sorted(data_vars))
    shift = np.stack([np.asarray(_quantile(quantiles, q_shift)[v], dtype=np.float64) for v in names])
    scale = np.stack([np.asarray(_quantile(quantiles, q_hi)[v], dtype=np.float64) -
                      np.asarray(_quantile(quantiles, q_lo)[v], dtype=np.float64) for v in names])
    if shift.ndim not in (1, 3) or scale.shape != shift.shape:
        raise ValueError("quantiles must be scalars or [H, W] fields per variable")
    return shift.astype(np.float32), scale.astype(np.float32)


def _device(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.C2WError("no CUDA device: climate2weather_b200 has no CPU path")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def normalize_pack(ds: Mapping[str, np.ndarray], quantiles: Mapping, mode: str, data_vars: Sequence[str] = None,
                   device=None) -> torch.Tensor:
    """`normalize_ds` + `ds_to_sorted_np` + trajectory packing: {var: [L, H, W]} -> device fp32 [L, H, W, C]
    (variables sorted).  `ds` may also be a stacked array/tensor [C, L, H, W] in sorted-variable order."""
    lib = _lib.load()
    dev = _device(device)
    names = list(sorted(data_vars if data_vars is not None else ds.keys()))
    shift, scale = coefficients(quantiles, names, mode)
    if isinstance(ds, Mapping):
        src = torch.from_numpy(np.stack([np.asarray(ds[v], dtype=np.float32) for v in names]))
    else:
        src = torch.as_tensor(ds, dtype=torch.float32)
    C, L, H, W = src.shape
    if C != len(names):
        raise ValueError(f"{C} stacked variables for {len(names)} names")
    field = int(shift.ndim == 3)
    if field and shift.shape[1:] != (H, W):
        raise ValueError(f"quantile fields {shift.shape[1:]} do not match the grid {(H, W)}")
    with torch.cuda.device(dev):
        src = src.to(dev, non_blocking=True).contiguous()
        sh, sc = torch.from_numpy(shift).to(dev), torch.from_numpy(scale).to(dev)
        out = torch.empty(L, H, W, C, dtype=torch.float32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.c2w_normalize_pack(src.data_ptr(), out.data_ptr(), L, C, H * W, 1, sh.data_ptr(), sc.data_ptr(),
                                          field, st), "c2w_normalize_pack")
    return out


def unpack_unnormalize(x: torch.Tensor, quantiles: Mapping, mode: str, data_vars: Sequence[str]) -> Dict[str, np.ndarray]:
    """trajectory unpacking + `np_to_ds` + `unnormalize_ds`: device fp32 [L, H, W, C] -> {var: float32 [L, H, W]} on the
    host, in physical units."""
    lib = _lib.load()
    if not x.is_cuda:
        raise _lib.C2WError("unpack_unnormalize expects the device trajectory [L, H, W, C]; there is no CPU path")
    names = list(sorted(data_vars))
    L, H, W, C = x.shape
    if C != len(names):
        raise ValueError(f"trajectory has {C} variables, {len(names)} names given")
    shift, scale = coefficients(quantiles, names, mode)
    dev = x.device
    with torch.cuda.device(dev):
        xs = x.to(torch.float32).contiguous()
        sh, sc = torch.from_numpy(shift).to(dev), torch.from_numpy(scale).to(dev)
        out = torch.empty(C, L, H, W, dtype=torch.float32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.c2w_unpack_unnormalize(xs.data_ptr(), out.data_ptr(), L, C, H * W, 1, sh.data_ptr(),
                                              sc.data_ptr(), int(shift.ndim == 3), st), "c2w_unpack_unnormalize")
        host = out.cpu().numpy()
    return {v: host[i] for i, v in enumerate(names)}
