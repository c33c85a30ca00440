"""Config schema of the acoustic model — the drop-in mirror of the reference's
``fs2/config/__init__.py:31-317``.

Only the field names, nesting and defaults are the contract (the model reads
``config.model.encoder.layers`` and friends in ``__init__`` only).  The reference
builds these classes on top of un-vendored ``everyvoice.config`` base classes
(``ConfigModel``, ``BaseTrainingConfig``, ``PreprocessingConfig``, ``TextConfig``);
those bases are re-declared here as plain pydantic models carrying just the fields
the hot path reads, so the package stands alone.
"""
from __future__ import annotations

from enum import Enum
from pathlib import Path
from typing import Annotated, Any, Optional, Union

from annotated_types import Ge
from pydantic import BaseModel, ConfigDict, Field, field_serializer, model_validator

LATEST_VERSION: str = "1.1"  # reference config/__init__.py:28


class ConfigModel(BaseModel):
    model_config = ConfigDict(extra="forbid", use_enum_values=False)


class TargetTrainingTextRepresentationLevel(str, Enum):
    # everyvoice.config.type_definitions (un-vendored); values used at model.py:74-77
    characters = "characters"
    phones = "phones"
    phonological_features = "phonological_features"


N_PHONOLOGICAL_FEATURES = 39  # everyvoice.text.features


class ConformerConfig(ConfigModel):  # config/__init__.py:31-48
    layers: int = 4
    heads: int = 2
    input_dim: int = 256
    feedforward_dim: int = 1024
    conv_kernel_size: int = 9
    dropout: float = 0.2


class VarianceLevelEnum(str, Enum):
    phone = "phone"
    frame = "frame"


class VarianceLossEnum(str, Enum):
    mse = "mse"
    mae = "mae"


class VariancePredictorBase(ConfigModel):  # config/__init__.py:67-93
    loss: VarianceLossEnum = VarianceLossEnum.mse
    n_layers: int = 5
    kernel_size: int = 3
    dropout: float = 0.5
    input_dim: int = 256
    n_bins: int = 256
    depthwise: bool = True

    @field_serializer("loss")
    def convert_loss_enum(self, loss: VarianceLossEnum):
        return loss.value


class VariancePredictorConfig(VariancePredictorBase):  # :96-105
    level: VarianceLevelEnum = VarianceLevelEnum.phone

    @field_serializer("level")
    def convert_level_enum(self, level: VarianceLevelEnum):
        return level.value


class VariancePredictors(ConfigModel):  # :108-120
    energy: VariancePredictorConfig = Field(default_factory=VariancePredictorConfig)
    duration: VariancePredictorBase = Field(default_factory=VariancePredictorBase)
    pitch: VariancePredictorConfig = Field(default_factory=VariancePredictorConfig)


class FastSpeech2ModelConfig(ConfigModel):  # :123-175
    encoder: ConformerConfig = Field(default_factory=ConformerConfig)
    decoder: ConformerConfig = Field(default_factory=ConformerConfig)
    variance_predictors: VariancePredictors = Field(default_factory=VariancePredictors)
    target_text_representation_level: TargetTrainingTextRepresentationLevel = (
        TargetTrainingTextRepresentationLevel.characters
    )
    learn_alignment: bool = True
    use_global_style_token_module: bool = False
    max_length: int = 1000
    mel_loss: VarianceLossEnum = VarianceLossEnum.mse
    use_postnet: bool = True
    multilingual: bool = False
    multispeaker: bool = False

    @field_serializer("mel_loss")
    def convert_mel_loss_enum(self, mel_loss: VarianceLossEnum):
        return mel_loss.value

    @field_serializer("target_text_representation_level")
    def convert_training_enum(self, v: TargetTrainingTextRepresentationLevel):
        return v.value


class NoamOptimizer(ConfigModel):
    # everyvoice.config.shared_types.NoamOptimizer (un-vendored); defaults as used at
    # config/__init__.py:198-203 and tests/data/config/everyvoice-text-to-spec.yaml:30-36
    learning_rate: float = 1e-3
    eps: float = 1e-8
    weight_decay: float = 1e-6
    betas: tuple[float, float] = (0.9, 0.999)
    name: str = "noam"
    warmup_steps: int = 1000


class FastSpeech2TrainingConfig(ConfigModel):  # :193-243 (+ the BaseTrainingConfig fields read here)
    model_config = ConfigDict(extra="allow")
    batch_size: int = 16
    max_epochs: int = 1000
    max_steps: int = 100000
    use_weighted_sampler: bool = False
    optimizer: NoamOptimizer = Field(default_factory=NoamOptimizer)
    vocoder_path: Union[Path, None] = None
    mel_loss_weight: float = 1.0
    postnet_loss_weight: float = 1.0
    pitch_loss_weight: float = 0.1
    energy_loss_weight: float = 0.1
    duration_loss_weight: float = 0.1
    attn_ctc_loss_weight: float = 0.1
    attn_bin_loss_weight: float = 0.1
    attn_bin_loss_warmup_epochs: Annotated[int, Ge(1)] = 100


class AudioConfig(ConfigModel):
    model_config = ConfigDict(extra="allow")
    n_mels: int = 80
    input_sampling_rate: int = 22050
    output_sampling_rate: int = 22050


class PreprocessingConfig(ConfigModel):
    model_config = ConfigDict(extra="allow")
    audio: AudioConfig = Field(default_factory=AudioConfig)
    save_dir: Path = Path("./preprocessed")


class TextConfig(ConfigModel):
    """Just enough of everyvoice's TextConfig to size the symbol table: the pad
    symbol is id 0 (model.py:83-89) followed by the sorted union of ``symbols``."""

    model_config = ConfigDict(extra="allow")
    symbols: dict[str, list[str]] = Field(
        default_factory=lambda: {"letters": [chr(ord("a") + i) for i in range(26)]}
    )


class FastSpeech2Config(ConfigModel):  # :246-317
    model_config = ConfigDict(extra="allow")
    VERSION: str = LATEST_VERSION
    model: FastSpeech2ModelConfig = Field(default_factory=FastSpeech2ModelConfig)
    training: FastSpeech2TrainingConfig = Field(default_factory=FastSpeech2TrainingConfig)
    preprocessing: PreprocessingConfig = Field(default_factory=PreprocessingConfig)
    text: TextConfig = Field(default_factory=TextConfig)

    @model_validator(mode="before")
    @classmethod
    def check_and_upgrade_checkpoint(cls, data: Any) -> Any:
        """Same version gate as the reference (:299-317): newer → ValueError, <1.0 → 1.0."""
        from packaging.version import Version

        if not isinstance(data, dict):
            return data
        ckpt_version = Version(data.get("VERSION", "0.0"))
        if ckpt_version > Version(LATEST_VERSION):
            raise ValueError(
                "Your config was created with a newer version of EveryVoice, please update your software."
            )
        if ckpt_version < Version("1.0"):
            data["VERSION"] = "1.0"
        return data

    @staticmethod
    def load_config_from_path(path: Path) -> "FastSpeech2Config":
        import json

        import yaml

        text = Path(path).read_text()
        data = json.loads(text) if str(path).endswith(".json") else yaml.safe_load(text)
        return FastSpeech2Config(**data)

    def model_checkpoint_dump(self) -> dict:
        return self.model_dump(mode="json")
