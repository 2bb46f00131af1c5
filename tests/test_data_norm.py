"""N4: normalisation fused with the layout change (data/pipeline.py:183-272).

CPU: the oracle restatement against outputs of the REFERENCE's own functions (tests/golden/data_norm.npz, written by
tests/golden/make_golden_data.py running data/pipeline.py on an xarray stand-in: all five modes, scalar and per-grid-
point quantiles, both orderings, the unnormalise round trip), against hand-computed known answers, and the host-side
coefficient logic.  GPU (`-m gpu`): `c2w_normalize_pack` / `c2w_unpack_unnormalize` through the package's `data` module against the
oracle for all five modes, scalar and per-grid-point quantiles, plus the round trip and the kernel's HBM rate.
"""
import numpy as np
import pytest
import torch

from oracle import data_ref

VARS = ("vas", "psl", "uas", "tas")  # deliberately unsorted: the reference sorts (data/pipeline.py:255)
QS = (0.0, 0.01, 0.05, 0.25, 0.5, 0.75, 0.95, 0.99, 1.0)


def _problem(L=5, H=16, W=24, field=False, seed=0):
    rng = np.random.default_rng(seed)
    scale = {"psl": 1.0e3, "tas": 15.0, "uas": 6.0, "vas": 5.0}
    base = {"psl": 1.0e5, "tas": 280.0, "uas": 0.5, "vas": -0.3}
    ds = {v: (base[v] + scale[v] * rng.standard_normal((L, H, W))).astype(np.float32) for v in VARS}
    quantiles = {}
    for q in QS:
        quantiles[q] = {}
        for v in VARS:
            val = base[v] + scale[v] * (4.0 * q - 2.0)
            quantiles[q][v] = (val + 0.01 * scale[v] * rng.standard_normal((H, W))) if field else np.float64(val)
    return ds, quantiles


# ---------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("tag", ["scalar", "field"])
def test_oracle_matches_reference_functions(golden_dir, tag):
    """oracle/data_ref.py against normalize_ds / ds_to_sorted_np / np_to_ds / unnormalize_ds of the reference itself
    (data/pipeline.py:183-272): float64 results equal to 1e-12 (same two numpy operations per mode)."""
    g = np.load(golden_dir / "data_norm.npz")
    qs = [float(q) for q in g["qs"]]
    names = [str(v) for v in g["vars"]]
    ds = {v: g[f"{tag}::data::{v}"] for v in names}
    quantiles = {q: {v: g[f"{tag}::quant::{v}"][i] for v in names} for i, q in enumerate(qs)}
    for mode in (str(m) for m in g["modes"]):
        n = data_ref.normalize_ds(ds, quantiles, mode)
        lchw = data_ref.ds_to_sorted_np(n, names)
        assert lchw.dtype == np.float64 and lchw.shape == g[f"{tag}::{mode}::normalized_lchw"].shape
        assert np.allclose(lchw, g[f"{tag}::{mode}::normalized_lchw"], rtol=1e-12, atol=0)
        assert np.allclose(data_ref.ds_to_sorted_np(n, names, ordering="CLHW"), g[f"{tag}::{mode}::normalized_clhw"],
                           rtol=1e-12, atol=0)
        back = data_ref.unnormalize_ds(data_ref.np_to_ds(lchw, names), quantiles, mode)
        assert np.allclose(data_ref.ds_to_sorted_np(back, names), g[f"{tag}::{mode}::roundtrip_lchw"], rtol=1e-12, atol=0)
    # variables come out sorted (data/pipeline.py:255): psl, tas, uas, vas
    assert np.array_equal(data_ref.ds_to_sorted_np(ds, names)[:, 0], ds["psl"])


def test_oracle_known_answers():
    ds = {"a": np.array([[[2.0, 4.0]]]), "b": np.array([[[10.0, 30.0]]])}
    quantiles = {q: {"a": 0.0, "b": 0.0} for q in QS}
    quantiles[0.0] = {"a": 2.0, "b": 10.0}
    quantiles[1.0] = {"a": 4.0, "b": 30.0}
    quantiles[0.25] = {"a": 2.5, "b": 15.0}
    quantiles[0.5] = {"a": 3.0, "b": 20.0}
    quantiles[0.75] = {"a": 3.5, "b": 25.0}
    quantiles[0.05] = {"a": 2.1, "b": 11.0}
    quantiles[0.95] = {"a": 3.9, "b": 29.0}
    quantiles[0.01] = {"a": 2.02, "b": 10.2}
    quantiles[0.99] = {"a": 3.98, "b": 29.8}
    n = data_ref.normalize_ds(ds, quantiles, "minmax")
    assert np.allclose(n["a"], [[[0.0, 1.0]]]) and np.allclose(n["b"], [[[0.0, 1.0]]])
    n = data_ref.normalize_ds(ds, quantiles, "robust")  # (x - median) / (q75 - q25)
    assert np.allclose(n["a"], [[[-1.0, 1.0]]]) and np.allclose(n["b"], [[[-1.0, 1.0]]])
    n = data_ref.normalize_ds(ds, quantiles, "robust95")  # (x - median) / (q95 - q05)
    assert np.allclose(n["a"], [[[-1.0 / 1.8, 1.0 / 1.8]]]) and np.allclose(n["b"], [[[-10.0 / 18, 10.0 / 18]]])
    n = data_ref.normalize_ds(ds, quantiles, "quant95")  # (x - q05) / (q95 - q05)
    assert np.allclose(n["a"], [[[-0.1 / 1.8, 1.9 / 1.8]]])
    n = data_ref.normalize_ds(ds, quantiles, "quant99")  # (x - q01) / (q99 - q01)
    assert np.allclose(n["b"], [[[-0.2 / 19.6, 19.8 / 19.6]]])
    with pytest.raises(ValueError, match="Invalid mode"):
        data_ref.normalize_ds(ds, quantiles, "zscore")
    for mode in data_ref.MODES:
        back = data_ref.unnormalize_ds(data_ref.normalize_ds(ds, quantiles, mode), quantiles, mode)
        assert np.allclose(back["a"], ds["a"]) and np.allclose(back["b"], ds["b"])


def test_oracle_layout_and_host_coefficients():
    from climate2weather_b200 import data as c2w_data

    ds, quantiles = _problem()
    arr = data_ref.ds_to_sorted_np(ds, VARS)
    assert arr.shape == (5, 4, 16, 24) and np.array_equal(arr[:, 0], ds["psl"]) and np.array_equal(arr[:, 3], ds["vas"])
    assert data_ref.ds_to_sorted_np(ds, VARS, "CLHW").shape == (4, 5, 16, 24)
    back = data_ref.np_to_ds(arr, VARS)
    assert all(np.array_equal(back[v], ds[v]) for v in VARS)
    assert c2w_data.NORM_MODES == data_ref.MODES
    for mode, (qs, (ql, qh)) in data_ref.MODES.items():
        shift, scale = c2w_data.coefficients(quantiles, VARS, mode)
        assert shift.shape == (4,) and shift.dtype == np.float32
        for i, v in enumerate(sorted(VARS)):
            assert shift[i] == np.float32(quantiles[qs][v]) and scale[i] == np.float32(quantiles[qh][v] - quantiles[ql][v])
    _, qf = _problem(field=True)
    shift, scale = c2w_data.coefficients(qf, VARS, "quant99")
    assert shift.shape == (4, 16, 24) and scale.shape == (4, 16, 24)
    with pytest.raises(ValueError, match="Invalid mode"):
        c2w_data.coefficients(quantiles, VARS, "zscore")
    if not torch.cuda.is_available():  # no CPU path
        from climate2weather_b200 import _lib
        with pytest.raises((_lib.C2WError, OSError)):
            c2w_data.normalize_pack(ds, quantiles, "minmax")


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("field", [False, True])
@pytest.mark.parametrize("mode", sorted(data_ref.MODES))
def test_normalize_pack_and_back_vs_oracle(mode, field):
    from climate2weather_b200 import data as c2w_data

    ds, quantiles = _problem(L=7, H=32, W=40, field=field, seed=3)
    want = data_ref.ds_to_sorted_np(data_ref.normalize_ds(ds, quantiles, mode), VARS)  # float64 [L, C, H, W]
    x = c2w_data.normalize_pack(ds, quantiles, mode, device="cuda:0")
    assert x.shape == (7, 32, 40, 4) and x.dtype == torch.float32 and x.is_cuda
    got = x.permute(0, 3, 1, 2).cpu().numpy()
    # fp32 subtraction of ~1e5-sized pressures against float64: 1 ulp(1e5) / scale = 8e-3 / 4e3
    assert np.max(np.abs(got - want)) <= 4e-6 * max(1.0, np.max(np.abs(want)))
    # and back: physical units per variable, against the float32 input
    back = c2w_data.unpack_unnormalize(x, quantiles, mode, VARS)
    for v in VARS:
        assert back[v].shape == ds[v].shape and back[v].dtype == np.float32
        assert np.max(np.abs(back[v] - ds[v])) <= 2e-6 * np.max(np.abs(ds[v]))
    # the oracle's un-normalisation of the oracle's normalisation is what the device pass reproduces
    want_back = data_ref.unnormalize_ds(data_ref.np_to_ds(want, VARS), quantiles, mode)
    for v in VARS:
        assert np.max(np.abs(back[v] - want_back[v])) <= 2e-6 * np.max(np.abs(want_back[v]))


@pytest.mark.gpu
def test_normalize_pack_other_layouts_and_rate():
    """Sorted-numpy source order (clhw = 0), a channel count other than 4, and the HBM rate at a year of frames."""
    import ctypes

    from climate2weather_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    g = torch.Generator().manual_seed(2)
    for C in (3, 4):
        L, H, W = 6, 8, 24
        src = torch.randn(L, C, H, W, generator=g).to(dev)
        shift, scale = torch.randn(C, generator=g).to(dev), (torch.rand(C, generator=g) + 0.5).to(dev)
        out = torch.empty(L, H, W, C, device=dev)
        _lib.check(lib.c2w_normalize_pack(src.data_ptr(), out.data_ptr(), L, C, H * W, 0, shift.data_ptr(),
                                          scale.data_ptr(), 0, st), "c2w_normalize_pack")
        want = ((src - shift.view(1, C, 1, 1)) / scale.view(1, C, 1, 1)).permute(0, 2, 3, 1)
        assert torch.equal(out, want)  # same IEEE fp32 operations
        back = torch.empty_like(src)
        _lib.check(lib.c2w_unpack_unnormalize(out.data_ptr(), back.data_ptr(), L, C, H * W, 0, shift.data_ptr(),
                                              scale.data_ptr(), 0, st), "c2w_unpack_unnormalize")
        assert torch.equal(back, (out * scale + shift).permute(0, 3, 1, 2))
    # rate: 2048 frames of 4 x 128 x 128 (0.5 GiB in, 0.5 GiB out)
    L, C, H, W = 2048, 4, 128, 128
    src = torch.randn(C, L, H, W, device=dev)
    out = torch.empty(L, H, W, C, device=dev)
    shift, scale = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        lib.c2w_normalize_pack(src.data_ptr(), out.data_ptr(), L, C, H * W, 1, shift.data_ptr(), scale.data_ptr(), 0, st)
    e0.record()
    for _ in range(5):
        lib.c2w_normalize_pack(src.data_ptr(), out.data_ptr(), L, C, H * W, 1, shift.data_ptr(), scale.data_ptr(), 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = 2 * src.numel() * 4 / ms / 1e6
    print(f"\nc2w_normalize_pack: {L} frames, {ms:.3f} ms, {gbs:.0f} GB/s (read + write)")
    assert torch.equal(out, src.permute(1, 2, 3, 0))
    assert gbs > 2000  # HBM-bound pass; the measured copy peak of this pool is ~6500 GB/s
