"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float64 like xarray's promotion of float32 data against float64
quantiles) of the reference's normalisation and layout helpers, data/pipeline.py:183-272, on plain dicts of arrays
instead of xarray Datasets.  Only tests/ may import this module; the product path is the CUDA library.

Pinned: tests/golden/data_norm.npz holds outputs of the REFERENCE's own four functions (tests/golden/make_golden_data.py
runs data/pipeline.py unmodified on a minimal xarray stand-in — xarray itself is not installed in this image) for all
five modes, scalar and per-grid-point quantiles and both orderings; tests/test_data_norm.py checks this restatement
against them to 1e-12, next to hand-computed known answers.

  normalize_ds    data/pipeline.py:183-215     unnormalize_ds   data/pipeline.py:218-247
  ds_to_sorted_np data/pipeline.py:250-261     np_to_ds         data/pipeline.py:264-272
"""
import numpy as np

# mode -> (quantile subtracted, quantiles whose difference divides)     data/pipeline.py:189-213
MODES = {
    "minmax": (0.0, (0.0, 1.0)),
    "robust": (0.5, (0.25, 0.75)),
    "robust95": (0.5, (0.05, 0.95)),
    "quant95": (0.05, (0.05, 0.95)),
    "quant99": (0.01, (0.01, 0.99)),
}


def _coeff(quantiles, var, mode):
    if mode not in MODES:
        raise ValueError(f"Invalid mode: {mode}")  # data/pipeline.py:214-215
    q_shift, (q_lo, q_hi) = MODES[mode]
    shift = np.asarray(quantiles[q_shift][var], dtype=np.float64)
    scale = np.asarray(quantiles[q_hi][var], dtype=np.float64) - np.asarray(quantiles[q_lo][var], dtype=np.float64)
    return shift, scale


def normalize_ds(ds, quantiles, mode):
    """ds: {var: [L, H, W]}; quantiles: {q: {var: scalar or [H, W]}} -> {var: (x - shift) / scale}."""
    out = {}
    for v, x in ds.items():
        shift, scale = _coeff(quantiles, v, mode)
        out[v] = (np.asarray(x) - shift) / scale
    return out


def unnormalize_ds(ds, quantiles, mode):
    out = {}
    for v, x in ds.items():
        shift, scale = _coeff(quantiles, v, mode)
        out[v] = np.asarray(x) * scale + shift
    return out


def ds_to_sorted_np(ds, data_vars, ordering="LCHW"):
    assert ordering in ["LCHW", "CLHW"]
    data_vars = list(sorted(data_vars))
    return np.stack([ds[v] for v in data_vars], axis=0 if ordering == "CLHW" else 1)


def np_to_ds(np_arr, data_vars):
    data_vars = list(sorted(data_vars))
    assert np_arr.shape[1] == len(data_vars)
    return {v: np_arr[:, i] for i, v in enumerate(data_vars)}
