"""The three small schemas the hot path consumes — `InferenceControl` (the per-call pitch / energy / duration
multipliers), `StatsInfo` and `Stats` (dataset statistics the variance adaptor builds its bucket edges from).
Same class names, field names, defaults and validation as the reference's fs2/type_definitions_heavy.py:15-37,
declared here from field tables."""
from typing import Optional

from pydantic import ConfigDict, create_model

_REQUIRED = ...

# multipliers applied to the predicted (or given) variances at synthesis time; 1.0 = unchanged
InferenceControl = create_model(
    "InferenceControl",
    __config__=ConfigDict(arbitrary_types_allowed=True),
    **{name: (float, 1.0) for name in ("pitch", "energy", "duration")},
)

# raw and normalised range of one variance over the training set
StatsInfo = create_model(
    "StatsInfo",
    **{name: (float, _REQUIRED) for name in ("min", "max", "std", "mean", "norm_min", "norm_max")},
)

# pitch / energy are mandatory (the variance adaptor cannot be built without them); the length statistics are optional
Stats = create_model(
    "Stats",
    pitch=(StatsInfo, _REQUIRED),
    energy=(StatsInfo, _REQUIRED),
    **{name: (Optional[StatsInfo], None) for name in ("character_length", "phone_length", "arpabet_length")},
)

__all__ = ["InferenceControl", "StatsInfo", "Stats"]
