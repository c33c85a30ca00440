"""FastSpeech2 acoustic model — the drop-in for reference fs2/model.py:38-549.

Same class name, constructor, `forward(batch, control, inference) -> dict` (16 keys), step hooks,
optimizer configuration, checkpoint hooks and state-dict layout; everything between the batch dict
and the output dict runs in libfs2k (sm_100a) kernels.  `pytorch_lightning` is optional: when it is
installed the class is a real `LightningModule`, otherwise a minimal stand-in base is used.
"""
from __future__ import annotations

import sys
from typing import Optional

import torch
from torch import nn

from .. import autograd as ag
from .. import autograd_fns as fns
from .. import ops
from .._lib import require_device
from .config import N_PHONOLOGICAL_FEATURES, FastSpeech2Config, TargetTrainingTextRepresentationLevel
from .conformer import Conformer
from .gst.model import StyleEncoder
from .layers import PositionalEmbedding, PostNet
from .loss import FastSpeech2Loss
from .noam import NoamLR
from .type_definitions_heavy import InferenceControl, Stats
from .variance_adaptor import VarianceAdaptor, _log_error

try:
    import pytorch_lightning as pl  # type: ignore

    _Base = pl.LightningModule
except Exception:  # pragma: no cover - lightning is not installed in the build image

    class _Base(nn.Module):
        """The slice of LightningModule the model touches (model.py:70, :387-389, :524-528)."""

        def __init__(self, *a, **k):
            super().__init__()
            self.current_epoch = 0
            self.global_step = 0
            self.logger = None
            self.hparams = {}
            self.logged: dict = {}

        def save_hyperparameters(self, *a, **k):
            pass

        def log_dict(self, d, *a, **k):
            self.logged.update(d)


LookupTable = dict
DEFAULT_LANG2ID: LookupTable = {}
DEFAULT_SPEAKER2ID: LookupTable = {}


class TextProcessor:
    """Stand-in for everyvoice.text.text_processor.TextProcessor: only what model.py:64,80-89 reads —
    the symbol list (pad symbol first, id 0) and `encode_text`."""

    _pad_symbol = "\x80"

    def __init__(self, text_config):
        syms = []
        for v in text_config.symbols.values():
            syms.extend(v)
        self.symbols = [self._pad_symbol] + sorted(set(syms))

    def encode_text(self, text):
        return [self.symbols.index(text)]


class FastSpeech2(_Base):
    _VERSION: str = "1.2"

    def __init__(self, config: dict | FastSpeech2Config, stats: Optional[dict | Stats] = None,
                 lang2id: LookupTable = DEFAULT_LANG2ID, speaker2id: LookupTable = DEFAULT_SPEAKER2ID):
        super().__init__()
        if not isinstance(config, FastSpeech2Config):
            from pydantic import ValidationError

            try:
                config = FastSpeech2Config(**config)
            except ValidationError as e:
                _log_error(str(e))
                raise TypeError(
                    "Unable to load config.  Possible causes: is it really a FastSpeech2Config? or the correct version?"
                ) from e
        if stats is not None and not isinstance(stats, Stats):
            stats = Stats(**stats)
        self.config = config
        self.batch_size = config.training.batch_size
        self.text_processor = TextProcessor(config.text)
        self.lang2id = lang2id
        self.speaker2id = speaker2id
        self.stats = stats
        self.save_hyperparameters(ignore=[])
        self.loss = FastSpeech2Loss(config=config)
        m = self.config.model
        self.text_input_layer: nn.Linear | nn.Embedding
        if m.target_text_representation_level == TargetTrainingTextRepresentationLevel.phonological_features:
            self.text_input_layer = nn.Linear(N_PHONOLOGICAL_FEATURES, m.encoder.input_dim, bias=False)
        else:
            self.text_input_layer = nn.Embedding(
                len(self.text_processor.symbols), m.encoder.input_dim,
                padding_idx=self.text_processor.encode_text(self.text_processor._pad_symbol)[0],
            )
        self.position_embedding = PositionalEmbedding(m.encoder.input_dim)
        if m.use_global_style_token_module:
            self.gst = StyleEncoder(idim=self.config.preprocessing.audio.n_mels)
        self.encoder = Conformer(input_dim=m.encoder.input_dim, num_heads=m.encoder.heads,
                                 ffn_dim=m.encoder.feedforward_dim, num_layers=m.encoder.layers,
                                 depthwise_conv_kernel_size=m.encoder.conv_kernel_size, dropout=m.encoder.dropout)
        if self.stats is None:
            _log_error(
                "Your model doesn't have a value for self.stats either because the file is missing or the checkpoint "
                "didn't save them. We cannot initialize the variance adaptors without variance predictor statistics."
            )
            self.variance_adaptor = None
        else:
            self.variance_adaptor = VarianceAdaptor(self.config, self.stats)
        self.decoder = Conformer(input_dim=m.decoder.input_dim, num_heads=m.decoder.heads,
                                 ffn_dim=m.decoder.feedforward_dim, num_layers=m.decoder.layers,
                                 depthwise_conv_kernel_size=m.decoder.conv_kernel_size, dropout=m.decoder.dropout)
        self.mel_linear = nn.Linear(m.decoder.input_dim, self.config.preprocessing.audio.n_mels)
        if m.use_postnet:
            self.postnet = PostNet(n_mel_channels=self.config.preprocessing.audio.n_mels)
            self.output_key = "postnet_output"
        else:
            self.output_key = "output"
        self.speaker_embedding = None
        if m.multispeaker:
            if len(self.speaker2id) == 0:
                _log_error("Your model is multispeaker but speaker2id LookupTable is empty")
                sys.exit(1)
            self.speaker_embedding = nn.Embedding(len(self.speaker2id), m.encoder.input_dim)
        self.language_embedding = None
        if m.multilingual:
            if len(self.lang2id) == 0:
                _log_error("Your model is multilingual but language2id LookupTable is empty")
                sys.exit(1)
            self.language_embedding = nn.Embedding(len(self.lang2id), m.encoder.input_dim)

    # ------------------------------------------------------------------------------------------
    def forward(self, batch, control=InferenceControl(), inference=False):
        """model.py:153-268.  All tensors of `batch` must live on the model's CUDA device."""
        require_device()
        m = self.config.model
        if "duration_control" in batch and batch["duration_control"][0]:
            control.duration = batch["duration_control"][0]
        teacher_forcing = bool(inference and batch["mel_lens"] is not None)
        src_lens = batch["src_lens"]
        msl = batch["max_src_len"]
        # a CUDA 0-d tensor would cost a host sync per call; the padded input width is the same number
        max_src_len = (batch["pfs"] if batch.get("text") is None else batch["text"]).shape[1] if (
            torch.is_tensor(msl) and msl.is_cuda) else int(msl)
        mel_lens = batch["mel_lens"]
        max_mel_len = batch["max_mel_len"]
        inv_freq = self.position_embedding.inv_freq
        src_mask = ops.lens_mask(src_lens, max_src_len)

        # text embedding + positional term (:183-190)
        if m.target_text_representation_level == TargetTrainingTextRepresentationLevel.phonological_features:
            text_inputs = batch["pfs"]
            inputs = ag.linear(text_inputs.contiguous(), self.text_input_layer.weight)
            x = fns.add_posenc(inputs, inv_freq, src_lens)
        else:
            text_inputs = batch["text"]
            inputs, x = fns.embed_posenc(text_inputs[:, :max_src_len], self.text_input_layer.weight, inv_freq, src_lens,
                                         self.text_input_layer.padding_idx)
        if ops.PRECISION == "bf16":
            ops.refresh_bf16_shadows(self.parameters())  # bf16 weight operands current (one check per forward)
        # encoder (:193)
        x, _ = self.encoder(x, src_lens)

        # style / speaker / language rows, broadcast over T (:196-213)
        rows = []
        if m.use_global_style_token_module:
            with ops.full_precision():
                if inference and torch.is_tensor(batch["mel_style_reference"]):
                    rows.append((self.gst(batch["mel_style_reference"]), None))
                elif inference and not teacher_forcing:
                    rows.append((self.gst.condition_on_gst_tokens(batch["text"].size(0)), None))
                else:
                    rows.append((self.gst(batch["mel"]), None))
        if m.multispeaker and self.speaker_embedding is not None:
            rows.append((self.speaker_embedding.weight, batch["speaker_id"]))
        if m.multilingual and self.language_embedding is not None:
            rows.append((self.language_embedding.weight, batch["language_id"]))
        if rows:
            x = fns.add_rows(x, rows)

        # aligner → MAS, predictors → bucketize / durations: discrete decisions, fp32-level arithmetic in every mode
        with ops.full_precision():
            va = self.variance_adaptor(inputs, x, batch, src_mask, control, inference=inference,
                                       teacher_forcing=teacher_forcing, inv_freq=inv_freq)
        tgt_mask = va["target_mask"]
        if inference and not teacher_forcing:  # :226-230
            mel_lens = ops.mask_lens(tgt_mask)
            max_mel_len = tgt_mask.shape[1]

        # decoder positional term (:233-238) — already added by the gather kernel for phone-level models
        if va["output_with_pos"] is not None:
            dec_in = va["output_with_pos"]
        else:
            dec_in = fns.add_posenc(va["output"], inv_freq, ops.mask_lens(tgt_mask))
        cut = getattr(self, "_backward_cut", None)
        if cut is not None and torch.is_grad_enabled() and dec_in.requires_grad:
            # data-parallel training: the backward is run in two phases around this tensor (graphs.GraphedTrainStep), so that the
            # decoder / PostNet gradients can be all-reduced while the variance adaptor and the encoder are still back-propagating
            dec_in = cut(dec_in)
        with ops.decoder_precision():  # no discrete decision follows: optional reduced-precision synthesis (ops.set_precision)
            x, _ = self.decoder(dec_in, mel_lens)
            output = ag.linear(x, self.mel_linear.weight, self.mel_linear.bias)
            postnet_output = None
            if m.use_postnet:
                postnet_output = fns.add(output, self.postnet(output))
        return {
            "output": output,
            "postnet_output": postnet_output,
            "src_mask": src_mask,
            "src_lens": src_lens,
            "tgt_mask": tgt_mask,
            "tgt_lens": mel_lens,
            "attn_logprob": va["attn_logprob"],
            "attn_soft": va["attn_soft"],
            "attn_hard": va["attn_hard"],
            "duration_prediction": va["duration_prediction"],
            "duration_target": va["duration_target"],
            "energy_prediction": va["energy_prediction"],
            "energy_target": va["energy_target"],
            "pitch_prediction": va["pitch_prediction"],
            "pitch_target": va["pitch_target"],
            "text_input": text_inputs,
        }

    # ------------------------------------------------------------------------------------------
    def check_and_upgrade_checkpoint(self, checkpoint):
        """Version gate of model.py:270-351 (type / newer-version errors, 0.0 → 1.0 bump).  Re-mapping the
        symbol table of pre-1.2 checkpoints needs everyvoice's symbol utilities and is rejected here."""
        from packaging.version import Version

        model_info = checkpoint.get("model_info", {"name": self.__class__.__name__, "version": "1.0"})
        ckpt_model_type = model_info.get("name", "MISSING_TYPE")
        if ckpt_model_type != self.__class__.__name__:
            raise TypeError(
                f"""Wrong model type ({ckpt_model_type}), we are expecting a '{self.__class__.__name__}' model""")
        ckpt_version = Version(model_info.get("version", "0.0"))
        if ckpt_version > Version(self._VERSION):
            raise ValueError("Your model was created with a newer version of EveryVoice, please update your software.")
        if ckpt_version < Version("1.0"):
            checkpoint.setdefault("model_info", model_info)["version"] = "1.0"
        if ckpt_version < Version("1.2"):
            raise ValueError(
                f"Checkpoint version {ckpt_version} predates the 1.2 symbol-table layout; upgrade it with the reference "
                "implementation (fs2/model.py:313-349) before loading it here.")
        return checkpoint

    def on_load_checkpoint(self, checkpoint):
        checkpoint = self.check_and_upgrade_checkpoint(checkpoint)
        self.config = FastSpeech2Config(**checkpoint["hyper_parameters"]["config"])
        if checkpoint["hyper_parameters"].get("stats") is not None:
            self.stats = Stats(**checkpoint["hyper_parameters"]["stats"])

    def on_save_checkpoint(self, checkpoint):
        checkpoint.setdefault("hyper_parameters", {})["config"] = self.config.model_checkpoint_dump()
        if self.stats is not None:
            checkpoint["hyper_parameters"]["stats"] = self.stats.model_dump(mode="json")
        checkpoint["model_info"] = {"name": self.__class__.__name__, "version": self._VERSION}

    # ------------------------------------------------------------------------------------------
    def enable_cuda_graphs(self, enabled: bool = True):
        """Replay teacher-forced synthesis (`predict_step` on batches that carry `mel_lens`) from CUDA graphs,
        one graph per (B, T, F) shape.  Turns the Σduration host check off (graphs cannot read back)."""
        from ..graphs import GraphedSynthesis

        self._graphed = GraphedSynthesis(self) if enabled else None
        if enabled and self.variance_adaptor is not None:
            self.variance_adaptor.validate_durations = False
        return self

    def optimization_step(self, batch, use_cuda_graph: bool = True):
        """training_step + backward + clip + optimizer.step + scheduler.step — what Lightning's fit loop does around
        `training_step` (fs2/model.py:384-390, fs2/cli/train.py:38) — as one call.  With `use_cuda_graph` the whole
        step of a repeated batch shape is one CUDA-graph replay (graphs.GraphedTrainStep).  Needs
        `configure_optimizers()` first.  Returns the dict of loss tensors (device scalars, no host read)."""
        runner = getattr(self, "_train_runner", None)
        if runner is None or runner.opt is not self.optimizer:
            from ..graphs import GraphedTrainStep

            # `train_graph_cache`: how many captured batch shapes are kept (LRU); raise it for shape-diverse streams
            runner = self._train_runner = GraphedTrainStep(self, self.optimizer, self.scheduler, max_graphs=int(getattr(self, "train_graph_cache", 16)))
        if use_cuda_graph:
            return runner(batch)
        dev = self.optimizer.flat_p.device
        runner._arm_seed_base()
        losses = runner._step_body({k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()})
        self.scheduler.step()
        return losses

    def predict_step(self, batch, batch_idx):
        graphed = getattr(self, "_graphed", None)
        if graphed is not None and batch.get("mel_lens") is not None and not self.training:
            return graphed(batch)
        with torch.no_grad():
            return self(batch, inference=True)

    def training_step(self, batch, batch_idx):
        output = self(batch)
        losses = self.loss(output, batch, self.current_epoch)
        self.log_dict({f"training/{k}_loss": v.item() for k, v in losses.items()}, prog_bar=True)
        return losses["total"]

    def validation_step(self, batch, batch_idx):
        """model.py:515-528 without the TensorBoard audio/figure logging (plotting is out of scope)."""
        output = self(batch)
        losses = self.loss(output, batch, self.current_epoch)
        self.log_dict({f"validation/{k}_loss": v.item() for k, v in losses.items()}, batch_size=self.batch_size,
                      sync_dist=True)

    def configure_optimizers(self):
        o = self.config.training.optimizer
        # Same update rule and hyper-parameters as the reference's torch.optim.AdamW (model.py:530-537), run as one
        # fused launch over a flat parameter buffer.  `fused_grad_clip` (e.g. 1.0, the value fs2/cli/train.py:38
        # hands to Lightning) folds clip_grad_norm_ into the same launch; leave it None when the trainer clips.
        from ..optim import FusedAdamW

        self.optimizer = FusedAdamW(self.parameters(), o.learning_rate, betas=tuple(o.betas), eps=o.eps,
                                    weight_decay=o.weight_decay, max_grad_norm=getattr(self, "fused_grad_clip", None))
        self.scheduler = NoamLR(self.optimizer, o.warmup_steps)
        return [self.optimizer], [{"scheduler": self.scheduler, "interval": "step"}]
