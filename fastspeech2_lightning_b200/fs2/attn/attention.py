"""ConvAttention — mirror of reference fs2/attn/attention.py:101-251 (RAD-TTS aligner)."""
import numpy as np
import torch
from torch import nn

from ... import functional as Fk


class ConvNorm(torch.nn.Module):
    """attention.py:22-56 (parameter container; the convs run in fs2k_gemm)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = torch.nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                    padding=padding, dilation=dilation, bias=bias)
        torch.nn.init.xavier_uniform_(self.conv.weight, gain=torch.nn.init.calculate_gain(w_init_gain))


class ConvAttention(torch.nn.Module):
    def __init__(self, n_mel_channels=80, n_speaker_dim=128, n_text_channels=512, n_att_channels=80,
                 temperature=1.0, n_mel_convs=2, align_query_enc_type="3xconv", use_query_proj=True):
        super().__init__()
        if align_query_enc_type != "3xconv" or not use_query_proj:
            raise NotImplementedError("FastSpeech2_lightning builds ConvAttention with '3xconv' query projection only")
        self.temperature = temperature
        self.att_scaling_factor = np.sqrt(n_att_channels)
        self.align_query_enc_type = align_query_enc_type
        self.use_query_proj = bool(use_query_proj)
        self.key_proj = nn.Sequential(
            ConvNorm(n_text_channels, n_text_channels * 2, kernel_size=3, bias=True, w_init_gain="relu"),
            torch.nn.ReLU(),
            ConvNorm(n_text_channels * 2, n_att_channels, kernel_size=1, bias=True),
        )
        self.query_proj = nn.Sequential(
            ConvNorm(n_mel_channels, n_mel_channels * 2, kernel_size=3, bias=True, w_init_gain="relu"),
            torch.nn.ReLU(),
            ConvNorm(n_mel_channels * 2, n_mel_channels, kernel_size=1, bias=True),
            torch.nn.ReLU(),
            ConvNorm(n_mel_channels, n_att_channels, kernel_size=1, bias=True),
        )

    def forward(self, queries, keys, query_lens, mask=None, key_lens=None, keys_encoded=None, attn_prior=None):
        """queries [B,C,T1] (mel), keys [B,C2,T2] (text) → (attn [B,1,T1,T2], attn_logprob).
        `mask` [B,T2,1] (True = padded key) is honoured through `key_lens`; when only the mask is given
        the key lengths are recovered from it."""
        q_blc = queries.transpose(1, 2).contiguous()
        k_blc = keys.transpose(1, 2).contiguous()
        return self.forward_blc(q_blc, k_blc, mask=mask, key_lens=key_lens, attn_prior=attn_prior)

    def forward_blc(self, q_blc, k_blc, mask=None, key_lens=None, attn_prior=None):
        """Channels-last entry point used by the model (avoids two transposes)."""
        if mask is None:
            lens = None
        elif key_lens is not None:
            lens = key_lens
        else:
            from ... import ops

            lens = ops.mask_lens(~mask.squeeze(-1))
        prior = attn_prior if torch.is_tensor(attn_prior) else None
        return Fk.conv_attention(self, q_blc, k_blc, lens, prior)
