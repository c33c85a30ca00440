"""Attention losses — mirror of reference fs2/attn/attention_loss.py:22-73."""
import torch


class AttentionCTCLoss(torch.nn.Module):
    """Forward-sum loss (:22-62).  Blank padding, key masking, log_softmax and the CTC recursion (targets
    1..key_len, input length query_len, mean reduction, zero_infinity) are one fused kernel pair
    (csrc/ctc.cu); lengths stay on the device, so the call never synchronises."""

    def __init__(self, blank_logprob=-1):
        super().__init__()
        self.blank_logprob = blank_logprob

    def forward(self, attn_logprob, in_lens, out_lens):
        from ... import autograd_fns as fns

        return fns.ctc_forward_sum(attn_logprob, in_lens, out_lens, self.blank_logprob)


class AttentionBinarizationLoss(torch.nn.Module):
    """−Σ log clamp(soft[hard==1], eps) / Σ hard (:65-73)."""

    def forward(self, hard_attention, soft_attention, eps=1e-12):
        from ... import autograd_fns as fns

        return fns.attn_bin_loss(hard_attention, soft_attention, eps)
