"""CPU oracle for the guided-sampling hot path of schmidtjonathan/Climate2Weather.

TEST INFRASTRUCTURE ONLY.  This package restates, in plain fp32 torch/numpy on the CPU, the
algorithm of the reference's hot path (model/nn.py, model/score.py, src/thor/score.py,
src/thor/pipelines.py, exp/downscaling.py:129-132).  It is the checker the CUDA path is
compared against; it is never the thing measured or shipped.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.  `climate2weather_b200/` never imports it and has no CPU fallback.

Pinning: the reference has no tests or golden vectors, so the oracle is pinned
against outputs of the reference's own code imported from /root/reference in the build
container (tests/golden/make_golden.py wrote tests/golden/*.npz; tests/test_oracle_golden.py
checks them).  One boundary stays "parity unpinned": `zuko.nn.LayerNorm` (zuko==1.0.1,
requirements.txt:33) is not vendored and not installed; its semantics are restated from the
published zuko 1.0.x source: (x - mean) / sqrt(var + eps) with torch.var's default (unbiased)
estimator, eps=1e-5, no affine parameters.
"""
