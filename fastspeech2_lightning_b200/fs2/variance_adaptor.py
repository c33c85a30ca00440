"""Variance adaptor — mirror of reference fs2/variance_adaptor.py (VariancePredictor :18-62,
LengthRegulator :65-81, VarianceAdaptor :84-412): same classes, constructor signatures, forward
signatures, output keys and state-dict names; the arithmetic runs in libfs2k kernels.
"""
from __future__ import annotations

import sys

import torch
from torch import nn

from .. import autograd_fns as fns
from .. import functional as Fk
from .. import ops
from .attn.attention import ConvAttention
from .config import FastSpeech2Config
from .layers import VarianceConvolutionLayer
from .type_definitions_heavy import InferenceControl, Stats, StatsInfo

try:  # the reference raises everyvoice's BadDataError (variance_adaptor.py:6,:300)
    from everyvoice.exceptions import BadDataError  # type: ignore
except Exception:  # everyvoice is not a dependency of this package

    class BadDataError(Exception):
        pass


def _log_error(msg: str) -> None:
    try:
        from loguru import logger

        logger.error(msg)
    except Exception:
        print(msg, file=sys.stderr)


class VariancePredictor(nn.Module):
    def __init__(self, input_dim: int, n_layers=5, n_channels=384, output_dim=1, kernel_size=5, dropout_rate=0.1,
                 depthwise: bool = False):
        super().__init__()
        if output_dim != 1:
            raise NotImplementedError("FastSpeech2_lightning only builds scalar predictors (output_dim=1)")
        self.conv = nn.ModuleList()
        self.kernel_size = kernel_size
        for idx in range(n_layers):
            self.conv.append(
                VarianceConvolutionLayer(
                    in_channels=input_dim if idx == 0 else n_channels, out_channels=n_channels,
                    kernel_size=kernel_size, dropout=dropout_rate, depthwise=depthwise,
                )
            )
        self.linear = nn.Linear(n_channels, output_dim)

    def forward(self, x, mask=None):
        """x [B,L,C] → [B,L] (·mask)  (variance_adaptor.py:55-62)."""
        return Fk.variance_predictor(x.contiguous(), mask, self, self.training)


class LengthRegulator(nn.Module):
    def forward(self, x, durations, max_length=None):
        """(out [B,F',D], mask [B,F']) with F' = min(max_b Σdur, max_length)  (variance_adaptor.py:65-81).
        One host read (the per-utterance totals) decides F'; everything else stays on the device."""
        out, _, mask, _ = self.expand(x, durations, max_length)
        return out, mask

    def expand(self, x, durations, max_length=None, inv_freq=None, known_width=None):
        cum, total = ops.lr_scan(durations)
        if known_width is not None:
            width = int(known_width)
            totals_host = None
        else:
            totals_host = total.cpu()
            width = int(totals_host.max()) if totals_host.numel() else 0
            if max_length is not None:
                width = min(width, int(max_length))
        out, out_pos, mask = fns.length_regulate(x, cum, total, width, inv_freq)
        return out, out_pos, mask, (cum, total, totals_host)


class VarianceAdaptor(nn.Module):
    """Variance Adaptor"""

    def __init__(self, config: FastSpeech2Config, stats: Stats):
        super().__init__()
        self.config = config
        self.stats = stats
        vp = self.config.model.variance_predictors
        d = self.config.model.encoder.input_dim

        def predictor(c):
            return VariancePredictor(input_dim=d, n_layers=c.n_layers, n_channels=c.input_dim, output_dim=1,
                                     kernel_size=c.kernel_size, dropout_rate=c.dropout, depthwise=c.depthwise)

        self.duration_predictor = predictor(vp.duration)
        self.length_regulator = LengthRegulator()
        self.pitch_predictor = predictor(vp.pitch)
        self.pitch_embedding = nn.Embedding(vp.pitch.n_bins, vp.pitch.input_dim)
        self.pitch_bins = nn.Parameter(
            torch.linspace(self.stats.pitch.norm_min, self.stats.pitch.norm_max, vp.pitch.n_bins - 1), requires_grad=False)
        self.energy_predictor = predictor(vp.energy)
        self.energy_embedding = nn.Embedding(vp.energy.n_bins, vp.energy.input_dim)
        self.energy_bins = nn.Parameter(
            torch.linspace(self.stats.energy.norm_min, self.stats.energy.norm_max, vp.energy.n_bins - 1), requires_grad=False)
        if self.config.model.learn_alignment:
            self.attention = ConvAttention(self.config.preprocessing.audio.n_mels, 0, d, use_query_proj=True,
                                           align_query_enc_type="3xconv")
        # Σ duration == mel_len sanity check needs a host read (BadDataError, :289-304); benchmarks may turn it off
        self.validate_durations = True
        self.last_bucket_ids: dict = {}

    # ------------------------------------------------------------------------------------------
    def binarize_attention(self, attn, in_lens, out_lens):
        """MAS on log(attn) per utterance window → dense 0/1 [B,1,F,T]; no gradient (:160-181)."""
        with torch.no_grad():
            _, _, hard = ops.mas(attn.detach(), in_lens, out_lens, take_log=True, dense=True)
        return hard

    def get_variance_embedding(self, x, target, mask, predictor, embedding, stats: StatsInfo, bins, control, inference):
        """(prediction, embed) like the reference (:183-205); the model itself uses the fused
        `_variance_embed_add`, which adds the gathered rows to x in the same pass."""
        prediction = predictor(x, mask)
        if not inference:
            ids = ops.bucketize(target.contiguous(), bins)
        else:
            prediction = fns.scale(prediction, control)
            ids = ops.bucketize(prediction.detach().contiguous(), bins)
        return prediction, fns.embedding_lookup(ids, embedding.weight)

    def _variance_embed_add(self, x, target, mask, predictor, embedding, bins, control, inference):
        prediction = predictor(x, mask)
        if not inference:
            y, ids = fns.bucketize_embed_add(target, 1.0, bins, embedding.weight, x)
        else:
            y, ids, prediction = fns.bucketize_embed_add(prediction, float(control), bins, embedding.weight, x, return_scaled=True)
        # kept for stage-wise parity checks (the ids are a discrete function of fp32 values)
        self.last_bucket_ids["pitch" if embedding is self.pitch_embedding else "energy"] = ids
        return prediction, y, ids

    def average_variance(self, var, durs):
        """phone-level mean of the non-zero frame values (:207-222)."""
        cum, _ = ops.lr_scan(durs)
        return ops.average_variance(var.contiguous(), cum)

    # ------------------------------------------------------------------------------------------
    def forward(self, text_emb, encoder_output, batch, src_mask, control=InferenceControl(), inference=False,
                teacher_forcing=False, inv_freq=None):
        cfgm = self.config.model
        x = encoder_output
        energy_target = batch["energy"] if not inference else None
        pitch_target = batch["pitch"] if not inference else None
        dur_in = batch["duration"]
        duration_target = dur_in if (torch.is_tensor(dur_in) or dur_in[0] is not None) else None
        max_target_len = batch["max_mel_len"]
        attn_logprob = attn_soft = attn_hard = None
        cum_total = None
        if (teacher_forcing or not inference) and cfgm.learn_alignment:  # :248-305
            attn_soft, attn_logprob = self.attention.forward_blc(
                batch["mel"], text_emb, mask=src_mask, key_lens=batch["src_lens"], attn_prior=batch["duration"])
            if not inference:  # the forward-sum loss only needs attn_logprob: start it now, off the critical path
                fns.ctc_forward_sum_prefetch(attn_logprob, batch["src_lens"], batch["mel_lens"])
            with torch.no_grad():
                path, duration_target, attn_hard = ops.mas(attn_soft.detach(), batch["src_lens"], batch["mel_lens"],
                                                           take_log=True, dense=True)
            if (pitch_target is not None and pitch_target.size(1) == text_emb.size(1)) or (
                energy_target is not None and energy_target.size(1) == text_emb.size(1)
            ):
                _log_error(
                    "Your pitch and/or energy targets are already averaged across phones, but when you are learning alignment with phone-level energy or pitch modelling, you must have an un-averaged target for these as the duration of phones changes during training. This should happen automatically if you re-run the preprocessing step for energy and pitch."
                )
                sys.exit(1)
            cum_total = ops.lr_scan(duration_target)
            if energy_target is not None and cfgm.variance_predictors.energy.level.value == "phone":
                energy_target = ops.average_variance(energy_target.contiguous(), cum_total[0])
            if pitch_target is not None and cfgm.variance_predictors.pitch.level.value == "phone":
                pitch_target = ops.average_variance(pitch_target.contiguous(), cum_total[0])
            if self.validate_durations:
                ok = (cum_total[1].cpu() == batch["mel_lens"].cpu().to(torch.int32))
                if not bool(ok.all()):
                    bad = [n for n, good in zip(batch["basename"], ok.tolist()) if not good]
                    raise BadDataError(f"Something failed with the following items, please check them for errors: {bad}")

        energy_prediction = pitch_prediction = None
        if cfgm.variance_predictors.energy.level.value == "phone":  # :309-329
            energy_prediction, x, _ = self._variance_embed_add(
                x, energy_target, src_mask, self.energy_predictor, self.energy_embedding, self.energy_bins,
                control.energy, inference)
        if cfgm.variance_predictors.pitch.level.value == "phone":  # :330-350
            pitch_prediction, x, _ = self._variance_embed_add(
                x, pitch_target, src_mask, self.pitch_predictor, self.pitch_embedding, self.pitch_bins,
                control.pitch, inference)

        log_duration_prediction = self.duration_predictor(x, mask=src_mask)  # :352
        if teacher_forcing or not inference:  # :354-358
            duration_rounded = duration_target
        else:  # :359-366
            duration_rounded = ops.round_durations(log_duration_prediction.detach(), control.duration)
        frame_level = "frame" in (cfgm.variance_predictors.energy.level.value, cfgm.variance_predictors.pitch.level.value)
        # With given durations (training / teacher forcing) Σdur == mel_lens, so the output width is the
        # batch's max_mel_len and no host read is needed; free-running synthesis reads the totals once.
        known_width = None
        if (teacher_forcing or not inference) and not self.validate_durations:
            mml = batch["max_mel_len"]
            if torch.is_tensor(mml) and mml.is_cuda:
                known_width = batch["mel"].shape[1] if batch.get("mel") is not None else int(mml)
            else:
                known_width = int(mml)
        x, x_pos, tgt_mask, scan = self.length_regulator.expand(
            x, duration_rounded, max_length=max_target_len, inv_freq=None if frame_level else inv_freq,
            known_width=known_width)

        if cfgm.variance_predictors.energy.level.value == "frame":  # :371-383
            energy_prediction, x, _ = self._variance_embed_add(
                x, energy_target, tgt_mask, self.energy_predictor, self.energy_embedding, self.energy_bins,
                control.energy, inference)
        if cfgm.variance_predictors.pitch.level.value == "frame":  # :385-397
            pitch_prediction, x, _ = self._variance_embed_add(
                x, pitch_target, tgt_mask, self.pitch_predictor, self.pitch_embedding, self.pitch_bins,
                control.pitch, inference)

        return {
            "output": x,
            "attn_logprob": attn_logprob,
            "attn_soft": attn_soft,
            "attn_hard": attn_hard,
            "duration_prediction": log_duration_prediction,
            "duration_target": duration_target,
            "pitch_prediction": pitch_prediction,
            "pitch_target": pitch_target,
            "energy_prediction": energy_prediction,
            "energy_target": energy_target,
            "duration_rounded": duration_rounded,
            "target_mask": tgt_mask,
            # extras (not in the reference dict): decoder input with the positional term already added
            # by the gather kernel, and the per-utterance frame totals (device, host)
            "output_with_pos": x_pos,
            "frame_totals": scan[1],
            "frame_totals_host": scan[2],
        }
