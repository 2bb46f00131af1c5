"""Makes the reference's import paths resolve to this package, so `exp/downscaling.py` / `training_loop.py` and
pickled snapshots (training_loop.py:250-266: {"ema": model.score.ScoreUNet, "pipeline": thor.pipelines.SDAPipeline})
work unchanged:

    import climate2weather_b200.compat as compat; compat.install()
    import thor.score, thor.pipelines, model.score        # -> this package
"""
from __future__ import annotations

import sys
import types

from . import model as _model
from . import pipelines as _pipelines
from . import score as _score


def install(force: bool = False) -> None:
    def mod(name: str, **attrs) -> types.ModuleType:
        if name in sys.modules and not force:
            m = sys.modules[name]
        else:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    thor = mod("thor")
    thor.score = mod("thor.score", AbstractScoreFunction=_score.AbstractScoreFunction,
                     DefaultScoreFunction=_score.DefaultScoreFunction,
                     BatchedScoreFunction=_score.BatchedScoreFunction, CoarseGrain=_score.CoarseGrain)
    thor.pipelines = mod("thor.pipelines", SDAPipeline=_pipelines.SDAPipeline)
    m = mod("model")
    m.score = mod("model.score", ScoreUNet=_model.ScoreUNet)
