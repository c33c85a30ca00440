"""FastSpeech2Loss — mirror of reference fs2/loss.py:9-126: same keys, weights and reductions
(means over the whole padded tensors after masking both operands)."""
from torch import nn

from .. import autograd_fns as fns
from .attn.attention_loss import AttentionBinarizationLoss, AttentionCTCLoss
from .config import FastSpeech2Config


class FastSpeech2Loss(nn.Module):
    def __init__(self, config: FastSpeech2Config):
        super().__init__()
        self.config = config
        self.attn_ctc_loss = AttentionCTCLoss()
        self.attn_bin_loss = AttentionBinarizationLoss()

    def forward(self, output, batch, current_epoch, frozen_components=None):
        m, t = self.config.model, self.config.training
        src_mask, tgt_mask = output["src_mask"], output["tgt_mask"]
        losses = {}
        for name, weight in (("pitch", t.pitch_loss_weight), ("energy", t.energy_loss_weight)):
            target = output[f"{name}_target"]
            if target is not None:
                vp = getattr(m.variance_predictors, name)
                mask = src_mask if vp.level.value == "phone" else tgt_mask
                losses[name] = fns.masked_loss(output[f"{name}_prediction"], target, mask, vp.loss.value, weight)
        # duration: log(duration_target + 1) vs predicted log-duration (:78-86)
        losses["duration"] = fns.masked_loss(output["duration_prediction"], output["duration_target"], src_mask,
                                             m.variance_predictors.duration.loss.value, t.duration_loss_weight,
                                             log1p_int_target=True)
        losses["spec"] = fns.masked_loss(output["output"], batch["mel"], tgt_mask, m.mel_loss.value, t.mel_loss_weight)
        if m.use_postnet:
            losses["postnet"] = fns.masked_loss(output["postnet_output"], batch["mel"], tgt_mask, m.mel_loss.value,
                                                t.postnet_loss_weight)
        if m.learn_alignment:
            ctc_loss = self.attn_ctc_loss(output["attn_logprob"], batch["src_lens"], batch["mel_lens"])
            losses["attn_ctc"] = ctc_loss * t.attn_ctc_loss_weight
            bin_loss_weight = min(current_epoch / t.attn_bin_loss_warmup_epochs, 1.0) * t.attn_bin_loss_weight
            losses["attn_bin"] = self.attn_bin_loss(output["attn_hard"], output["attn_soft"]) * bin_loss_weight
        losses["total"] = sum(losses.values())
        return losses
