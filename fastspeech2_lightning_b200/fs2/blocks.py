"""Mirror of reference fs2/blocks.py:4-87 — parameter containers + kernel-backed forwards.
All forwards take/return the reference's layouts ([B,C,L] for conv modules)."""
from torch import nn

from .. import autograd as ag


class DepthwiseSeparableConv1d(nn.Module):
    """depthwise k (groups=C, bias) → pointwise 1×1 (blocks.py:4-19).  forward: [B,C,L] → [B,Cout,L]."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int):
        super().__init__()
        self.model = nn.Sequential(
            nn.Conv1d(in_channels, in_channels, kernel_size, padding=(kernel_size - 1) // 2, groups=in_channels),
            nn.Conv1d(in_channels, out_channels, 1),
        )

    def forward(self, x):
        h = ag.dwconv(x.transpose(1, 2).contiguous(), self.model[0].weight, self.model[0].bias)
        h = ag.linear(h, self.model[1].weight.squeeze(-1), self.model[1].bias)
        return h.transpose(1, 2)


class ConvNorm(nn.Module):
    """xavier-initialised Conv1d (blocks.py:44-87).  forward: [B,Cin,L] → [B,Cout,L]."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear", transpose=False):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        if stride != 1 or dilation != 1 or padding != (kernel_size - 1) // 2:
            raise NotImplementedError("only stride 1, dilation 1, 'same' padding are used by the model")
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))
        self.transpose = transpose

    def forward(self, x):
        x_blc = x.contiguous() if self.transpose else x.transpose(1, 2).contiguous()
        y = ag.conv1d(x_blc, self.conv.weight, self.conv.bias)
        return y if self.transpose else y.transpose(1, 2)
