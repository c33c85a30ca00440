"""Attention losses — mirror of reference fs2/attn/attention_loss.py:22-73."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class AttentionCTCLoss(torch.nn.Module):
    """Forward-sum loss (:22-62).  The blank-padding, key masking and log_softmax are restated with the
    same tensor ops; the CTC recursion itself is torch's `ctc_loss` CUDA kernel — a library call,
    listed as the next component to replace."""

    def __init__(self, blank_logprob=-1):
        super().__init__()
        self.blank_logprob = blank_logprob

    def forward(self, attn_logprob, in_lens, out_lens):
        key_lens, query_lens = in_lens, out_lens
        max_key_len = attn_logprob.size(-1)
        lp = attn_logprob.squeeze(1).permute(1, 0, 2)
        lp = F.pad(input=lp, pad=(1, 0, 0, 0, 0, 0), value=self.blank_logprob)
        key_inds = torch.arange(max_key_len + 1, device=lp.device, dtype=torch.long)
        lp = lp.masked_fill(key_inds.view(1, 1, -1) > key_lens.view(1, -1, 1), -1e15)
        lp = F.log_softmax(lp, dim=-1)
        target_seqs = key_inds[1:].unsqueeze(0).repeat(key_lens.numel(), 1)
        return F.ctc_loss(lp, target_seqs, input_lengths=query_lens, target_lengths=key_lens, blank=0,
                          reduction="mean", zero_infinity=True)


class AttentionBinarizationLoss(torch.nn.Module):
    """−Σ log clamp(soft[hard==1], eps) / Σ hard (:65-73)."""

    def forward(self, hard_attention, soft_attention, eps=1e-12):
        from ... import autograd_fns as fns

        return fns.attn_bin_loss(hard_attention, soft_attention, eps)
