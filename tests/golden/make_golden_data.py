"""Generates tests/golden/data_norm.npz by running the REFERENCE's own normalisation / layout helpers
(data/pipeline.py:183-272: normalize_ds, unnormalize_ds, ds_to_sorted_np, np_to_ds) in the build container.

    python tests/golden/make_golden_data.py

`data/pipeline.py` imports xarray at module level and xarray is not installed here, so a minimal stand-in is registered
as `xarray` first: a Dataset is a dict of named float arrays with a `quantile` coordinate where present; `.sel(quantile=q)`
picks a slice, Dataset arithmetic is variable-wise numpy broadcasting with numpy's own type promotion (float32 data
against float64 quantiles -> float64, which is what xarray does too).  The reference functions themselves run
unmodified on it — only the container type is ours.  No reference source is copied.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


class DataArray:
    def __init__(self, values):
        self.values = np.asarray(values)


class Dataset:
    """dict of variables (+ an optional `quantile` axis 0 on every variable) with xarray's arithmetic semantics for the
    operations data/pipeline.py performs."""

    def __init__(self, data_vars=None, coords=None, quantiles=None):
        self.vars = {}
        for k, v in (data_vars or {}).items():
            if isinstance(v, tuple):  # (dims, array) as np_to_ds builds them
                v = v[1]
            self.vars[k] = np.asarray(v)
        self.coords = dict(coords or {})
        self.quantiles = None if quantiles is None else [float(q) for q in quantiles]
        for name in ("time", "rlat", "rlon"):
            if name in self.coords:
                setattr(self, name, self.coords[name])

    @property
    def dims(self):
        return {k: len(v) for k, v in self.coords.items()}

    def sel(self, quantile):
        i = self.quantiles.index(float(quantile))
        return Dataset({k: v[i] for k, v in self.vars.items()}, self.coords)

    def __getitem__(self, k):
        return DataArray(self.vars[k])

    def _bin(self, other, op):
        if isinstance(other, Dataset):
            return Dataset({k: op(v, other.vars[k]) for k, v in self.vars.items() if k in other.vars}, self.coords)
        return Dataset({k: op(v, other) for k, v in self.vars.items()}, self.coords)

    def __sub__(self, o):
        return self._bin(o, np.subtract)

    def __add__(self, o):
        return self._bin(o, np.add)

    def __mul__(self, o):
        return self._bin(o, np.multiply)

    def __truediv__(self, o):
        return self._bin(o, np.divide)


def load_reference_pipeline():
    xr = types.ModuleType("xarray")
    xr.Dataset = Dataset
    xr.load_dataset = xr.open_dataset = lambda path: (_ for _ in ()).throw(RuntimeError("no file I/O in the stand-in"))
    sys.modules["xarray"] = xr
    spec = importlib.util.spec_from_file_location("ref_data_pipeline", REF / "data/pipeline.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


QS = [0.0, 0.01, 0.05, 0.25, 0.5, 0.75, 0.95, 0.99, 1.0]
VARS = ["vas", "psl", "uas", "tas"]  # deliberately unsorted: ds_to_sorted_np / np_to_ds sort them
MODES = ["minmax", "robust", "robust95", "quant95", "quant99"]


def problem(field: bool):
    rng = np.random.default_rng(7 + int(field))
    L, H, W = 5, 6, 8
    data = {v: (rng.standard_normal((L, H, W)) * (3.0 + i) + 10.0 * i).astype(np.float32) for i, v in enumerate(VARS)}
    if field:  # per-grid-point quantiles [quantile, H, W]
        quant = {v: np.sort(rng.standard_normal((len(QS), H, W)) * 4.0 + 10.0 * i, axis=0) for i, v in enumerate(VARS)}
    else:      # scalar quantiles [quantile]
        quant = {v: np.sort(rng.standard_normal(len(QS)) * 4.0 + 10.0 * i) for i, v in enumerate(VARS)}
    return data, quant, (L, H, W)


def main():
    dp = load_reference_pipeline()
    out = {"qs": np.array(QS), "vars": np.array(VARS), "modes": np.array(MODES)}
    for field in (False, True):
        data, quant, (L, H, W) = problem(field)
        tag = "field" if field else "scalar"
        coords = dict(time=np.arange(L), rlat=np.arange(H), rlon=np.arange(W))
        ds = Dataset(data, coords)
        qds = Dataset(quant, coords, quantiles=QS)
        for v in VARS:
            out[f"{tag}::data::{v}"] = data[v]
            out[f"{tag}::quant::{v}"] = quant[v]
        for mode in MODES:
            nds = dp.normalize_ds(ds, qds, mode)
            arr = dp.ds_to_sorted_np(nds, VARS)            # [L, C, H, W], variables sorted
            out[f"{tag}::{mode}::normalized_lchw"] = arr
            out[f"{tag}::{mode}::normalized_clhw"] = dp.ds_to_sorted_np(nds, VARS, ordering="CLHW")
            back = dp.unnormalize_ds(dp.np_to_ds(arr, ds, VARS), qds, mode)
            out[f"{tag}::{mode}::roundtrip_lchw"] = dp.ds_to_sorted_np(back, VARS)
    np.savez_compressed(OUT / "data_norm.npz", **out)
    print("wrote data_norm.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
