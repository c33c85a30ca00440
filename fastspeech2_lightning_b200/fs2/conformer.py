"""Conformer encoder/decoder — the drop-in for `torchaudio.models.Conformer` as the reference
uses it (fs2/model.py:23,95-102,112-119,193,241).

Same constructor signature, same `forward(input[B,T,D], lengths[B]) -> (out, lengths)`, same
state-dict key names/shapes as torchaudio's `models/conformer.py` (so reference checkpoints load),
but the arithmetic runs in libfs2k kernels on channels-last `[B,L,D]` data: the sub-modules below
are parameter containers whose own `forward` is never called.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from .. import functional as Fk


class _ConvolutionModule(nn.Module):
    """Parameter layout of torchaudio conformer.py:18-88."""

    def __init__(self, input_dim: int, num_channels: int, depthwise_kernel_size: int, dropout: float = 0.0,
                 bias: bool = False, use_group_norm: bool = False) -> None:
        super().__init__()
        if (depthwise_kernel_size - 1) % 2 != 0:
            raise ValueError("depthwise_kernel_size must be odd to achieve 'SAME' padding.")
        if use_group_norm:
            raise NotImplementedError("use_group_norm is not used by FastSpeech2_lightning")
        self.layer_norm = nn.LayerNorm(input_dim)
        self.sequential = nn.Sequential(
            nn.Conv1d(input_dim, 2 * num_channels, 1, stride=1, padding=0, bias=bias),
            nn.GLU(dim=1),
            nn.Conv1d(num_channels, num_channels, depthwise_kernel_size, stride=1,
                      padding=(depthwise_kernel_size - 1) // 2, groups=num_channels, bias=bias),
            nn.BatchNorm1d(num_channels),
            nn.SiLU(),
            nn.Conv1d(num_channels, input_dim, kernel_size=1, stride=1, padding=0, bias=bias),
            nn.Dropout(dropout),
        )


class _FeedForwardModule(nn.Module):
    """Parameter layout of torchaudio conformer.py:91-119."""

    def __init__(self, input_dim: int, hidden_dim: int, dropout: float = 0.0) -> None:
        super().__init__()
        self.sequential = nn.Sequential(
            nn.LayerNorm(input_dim),
            nn.Linear(input_dim, hidden_dim, bias=True),
            nn.SiLU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, input_dim, bias=True),
            nn.Dropout(dropout),
        )


class ConformerLayer(nn.Module):
    """torchaudio conformer.py:122-212 (convolution_first=False, BatchNorm variant)."""

    def __init__(self, input_dim: int, ffn_dim: int, num_attention_heads: int, depthwise_conv_kernel_size: int,
                 dropout: float = 0.0, use_group_norm: bool = False, convolution_first: bool = False) -> None:
        super().__init__()
        if convolution_first:
            raise NotImplementedError("convolution_first is not used by FastSpeech2_lightning")
        self.ffn1 = _FeedForwardModule(input_dim, ffn_dim, dropout=dropout)
        self.self_attn_layer_norm = nn.LayerNorm(input_dim)
        self.self_attn = nn.MultiheadAttention(input_dim, num_attention_heads, dropout=dropout)
        self.self_attn_dropout = nn.Dropout(dropout)
        self.conv_module = _ConvolutionModule(
            input_dim=input_dim, num_channels=input_dim, depthwise_kernel_size=depthwise_conv_kernel_size,
            dropout=dropout, bias=True, use_group_norm=use_group_norm,
        )
        self.ffn2 = _FeedForwardModule(input_dim, ffn_dim, dropout=dropout)
        self.final_layer_norm = nn.LayerNorm(input_dim)
        self.num_heads = num_attention_heads
        self.dropout_p = dropout


class Conformer(nn.Module):
    def __init__(self, input_dim: int, num_heads: int, ffn_dim: int, num_layers: int,
                 depthwise_conv_kernel_size: int, dropout: float = 0.0, use_group_norm: bool = False,
                 convolution_first: bool = False):
        super().__init__()
        self.conformer_layers = nn.ModuleList(
            [
                ConformerLayer(input_dim, ffn_dim, num_heads, depthwise_conv_kernel_size, dropout=dropout,
                               use_group_norm=use_group_norm, convolution_first=convolution_first)
                for _ in range(num_layers)
            ]
        )

    def forward(self, input: torch.Tensor, lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """input [B,T,D] fp32 CUDA, lengths [B] → (output [B,T,D], lengths).  torchaudio conformer.py:273-293."""
        x = Fk.conformer_stack(input, lengths, self, self.training)
        return x, lengths
