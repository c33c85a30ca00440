"""Mirror of reference fs2/layers.py — PositionalEmbedding (:123-140), PostNet (:143-212),
VarianceConvolutionLayer (:20-48), Transpose (:11-17)."""
import torch
from torch import nn

from .. import autograd as ag
from .. import functional as Fk
from .. import ops
from .blocks import ConvNorm, DepthwiseSeparableConv1d


class Transpose(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        return self.module(x.transpose(1, 2)).transpose(1, 2)


class VarianceConvolutionLayer(nn.Module):
    """conv (depthwise-separable or plain) → ReLU → LayerNorm → Dropout on [B,L,C] (layers.py:20-48)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, dropout: float, depthwise: bool):
        super().__init__()
        conv = Transpose(
            DepthwiseSeparableConv1d(in_channels, out_channels, kernel_size)
            if depthwise
            else nn.Conv1d(in_channels, out_channels, kernel_size, padding=(kernel_size - 1) // 2)
        )
        self.layers = nn.Sequential(conv, nn.ReLU(), nn.LayerNorm(out_channels), nn.Dropout(dropout))

    def forward(self, x):
        conv, ln = self.layers[0].module, self.layers[2]
        p = self.layers[3].p if self.training else 0.0
        if hasattr(conv, "model"):
            h = ag.dwconv(x, conv.model[0].weight, conv.model[0].bias)
            h = ag.linear(h, conv.model[1].weight.squeeze(-1), conv.model[1].bias, act="relu")
        else:
            h = ag.conv1d(x, conv.weight, conv.bias, act="relu")
        return ag.layernorm(h, ln.weight, ln.bias, ln.eps, dropout=p)


class PositionalEmbedding(nn.Module):
    """FastPitch sinusoid: cat(sin(p·ω), cos(p·ω)), ω_i = 10000^(−2i/D) (layers.py:123-140)."""

    def __init__(self, embedding_dim):
        super().__init__()
        self.demb = embedding_dim
        inv_freq = 1 / (10000 ** (torch.arange(0.0, self.demb, 2.0) / self.demb))
        self.register_buffer("inv_freq", inv_freq)

    def forward(self, pos_seq, bsz=None):
        """pos_seq must be arange(L) (the only way the reference calls it, model.py:186-190,:233-238)."""
        L = pos_seq.shape[0]
        zeros = torch.zeros((1, L, self.demb), dtype=torch.float32, device=pos_seq.device)
        lens = torch.full((1,), L, dtype=torch.int32, device=pos_seq.device)
        pos_emb = ops.add_posenc(zeros, self.inv_freq, lens)
        return pos_emb.expand(bsz, -1, -1) if bsz is not None else pos_emb


class PostNet(nn.Module):
    """Five Conv1d(k=5) + BatchNorm1d blocks, tanh on the first four, dropout 0.5 in training
    (layers.py:143-212).  forward takes and returns [B,L,n_mel] like the reference."""

    def __init__(self, n_mel_channels=80, postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5):
        super().__init__()
        self.convolutions = nn.ModuleList()
        pad = int((postnet_kernel_size - 1) / 2)
        dims = [n_mel_channels] + [postnet_embedding_dim] * (postnet_n_convolutions - 1) + [n_mel_channels]
        for i in range(postnet_n_convolutions):
            gain = "tanh" if i < postnet_n_convolutions - 1 else "linear"
            self.convolutions.append(
                nn.Sequential(
                    ConvNorm(dims[i], dims[i + 1], kernel_size=postnet_kernel_size, stride=1, padding=pad, dilation=1,
                             w_init_gain=gain),
                    nn.BatchNorm1d(dims[i + 1]),
                )
            )
        # the reference hard-codes F.dropout(p=0.5, self.training) (layers.py:208-209); parity runs switch it off
        self.dropout_in_training = True

    def forward(self, x):
        return Fk.postnet(x.contiguous(), self, self.training)
