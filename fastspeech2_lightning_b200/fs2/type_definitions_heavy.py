"""InferenceControl / StatsInfo / Stats — the schemas of reference fs2/type_definitions_heavy.py:15-37 (field names, types and
defaults are the interface: plain pydantic models, as in the reference)."""
from typing import Optional

from pydantic import BaseModel, ConfigDict


class InferenceControl(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True)
    pitch: float = 1.0
    energy: float = 1.0
    duration: float = 1.0


class StatsInfo(BaseModel):
    min: float
    max: float
    std: float
    mean: float
    norm_min: float
    norm_max: float


class Stats(BaseModel):
    pitch: StatsInfo
    energy: StatsInfo
    character_length: Optional[StatsInfo] = None
    phone_length: Optional[StatsInfo] = None
    arpabet_length: Optional[StatsInfo] = None
