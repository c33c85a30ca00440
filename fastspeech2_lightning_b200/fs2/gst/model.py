"""Global style tokens — mirror of reference fs2/gst/model.py:14-280 / gst/attn.py:48-194
(parameter names kept so checkpoints load).

* `condition_on_gst_tokens` (reference-free synthesis, gst/model.py:77-85): a single key ⇒ the
  softmax is identically 1 ⇒ linear_out(linear_v(tanh(gst_embs[index]))), two libfs2k GEMMs.
* `forward(speech)` (reference encoder: 6×Conv2d s2 + BatchNorm2d + ReLU → GRU → 4-head token
  attention) runs on libfs2k kernels (csrc/gst.cu): synthesis / validation with
  folded BatchNorm in one fused conv kernel; training through autograd Functions over the raw conv, the
  BatchNorm batch-statistics kernels, conv dgrad / wgrad, a GRU with back-propagation through time and
  the token-attention backward.  There is no library (cuDNN / ATen) path and no CPU path: the torch layers below
  are parameter containers whose own forward is never called.
"""
from collections.abc import Sequence

import torch

from ... import autograd as ag
from ... import ops


def _fused_eval_path(module, *tensors) -> bool:
    """No gradient and eval mode: the fused kernels with folded BatchNorm; otherwise the autograd Functions over the
    training kernels (batch statistics, dgrad / wgrad, BPTT)."""
    return (not module.training) and not (torch.is_grad_enabled() and (
        any(t.requires_grad for t in tensors) or any(p.requires_grad for p in module.parameters())))


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{what} must be a CUDA tensor: the GST modules run on libfs2k kernels only (no CPU / library path)")


class MultiHeadedAttention(torch.nn.Module):
    def __init__(self, q_dim, k_dim, v_dim, n_head, n_feat, dropout_rate=0.0):
        super().__init__()
        assert n_feat % n_head == 0
        self.d_k = n_feat // n_head
        self.h = n_head
        self.linear_q = torch.nn.Linear(q_dim, n_feat)
        self.linear_k = torch.nn.Linear(k_dim, n_feat)
        self.linear_v = torch.nn.Linear(v_dim, n_feat)
        self.linear_out = torch.nn.Linear(n_feat, n_feat)
        self.dropout = torch.nn.Dropout(p=dropout_rate)

    def forward(self, query, key, value, mask=None):
        raise RuntimeError("parameter container of gst/attn.py:96-194: StyleTokenLayer.forward runs the kernels")


class ReferenceEncoder(torch.nn.Module):
    def __init__(self, idim=80, conv_layers: int = 6, conv_chans_list: Sequence[int] = (32, 32, 64, 64, 128, 128),
                 conv_kernel_size: int = 3, conv_stride: int = 2, gru_layers: int = 1, gru_units: int = 128):
        super().__init__()
        assert conv_kernel_size % 2 == 1, "kernel size must be odd."
        assert len(conv_chans_list) == conv_layers
        convs = []
        padding = (conv_kernel_size - 1) // 2
        for i in range(conv_layers):
            conv_in_chans = 1 if i == 0 else conv_chans_list[i - 1]
            conv_out_chans = conv_chans_list[i]
            convs += [
                torch.nn.Conv2d(conv_in_chans, conv_out_chans, kernel_size=conv_kernel_size, stride=conv_stride,
                                padding=padding, bias=False),
                torch.nn.BatchNorm2d(conv_out_chans),
                torch.nn.ReLU(inplace=True),
            ]
        self.convs = torch.nn.Sequential(*convs)
        gru_in_units = idim
        for _ in range(conv_layers):
            gru_in_units = (gru_in_units - conv_kernel_size + 2 * padding) // conv_stride + 1
        gru_in_units *= conv_out_chans
        self.gru = torch.nn.GRU(gru_in_units, gru_units, gru_layers, batch_first=True)

    def forward(self, speech: torch.Tensor) -> torch.Tensor:
        """gst/model.py:179-199.  CUDA inputs run on libfs2k kernels (fused eval path or autograd Functions)."""
        _require_cuda(speech, "speech")
        if _fused_eval_path(self, speech):
            x = speech.detach().contiguous().unsqueeze(-1)  # channels-last image [B, F, n_mels, 1]
            n = len(self.convs) // 3
            for i in range(n):
                conv, bn = self.convs[3 * i], self.convs[3 * i + 1]
                scale, shift = ops.bn_scale_shift(bn, None, False)
                w = conv.weight.detach().permute(2, 3, 1, 0).contiguous()  # [Co,Ci,3,3] → [3,3,Ci,Co]
                x = ops.conv2d_s2_bn_relu(x, w, scale, shift, cw_layout=(i == n - 1))
            hs = x.reshape(x.shape[0], x.shape[1], -1)  # [B, T', C·W']
            g = self.gru
            return ops.gru_last_hidden(hs, g.weight_ih_l0.detach(), g.weight_hh_l0.detach(), g.bias_ih_l0.detach(),
                                       g.bias_hh_l0.detach())
        from ... import autograd_fns as fns

        x = speech.contiguous().unsqueeze(-1)  # channels-last image [B, F, n_mels, 1]
        n = len(self.convs) // 3
        for i in range(n):
            x = fns.conv2d_s2_bn_relu(x, self.convs[3 * i], self.convs[3 * i + 1], self.training)
        hs = x.permute(0, 1, 3, 2).reshape(x.shape[0], x.shape[1], -1)  # [B,T',W',C] → [B,T',C·W'] (:195-197)
        return fns.gru_last_hidden(hs, self.gru)


class StyleTokenLayer(torch.nn.Module):
    def __init__(self, ref_embed_dim: int = 128, gst_tokens: int = 10, gst_token_dim: int = 256, gst_heads: int = 4,
                 dropout_rate: float = 0.0):
        super().__init__()
        self.register_parameter("gst_embs", torch.nn.Parameter(torch.randn(gst_tokens, gst_token_dim // gst_heads)))
        self.mha = MultiHeadedAttention(q_dim=ref_embed_dim, k_dim=gst_token_dim // gst_heads,
                                        v_dim=gst_token_dim // gst_heads, n_head=gst_heads, n_feat=gst_token_dim,
                                        dropout_rate=dropout_rate)

    def forward(self, ref_embs: torch.Tensor) -> torch.Tensor:
        _require_cuda(ref_embs, "ref_embs")
        if _fused_eval_path(self, ref_embs):
            m = self.mha
            tokens = ops.tanh(self.gst_embs.detach())
            q = ops.gemm(ref_embs.detach().contiguous(), m.linear_q.weight.detach(), m.linear_q.bias.detach())
            k = ops.gemm(tokens, m.linear_k.weight.detach(), m.linear_k.bias.detach())
            v = ops.gemm(tokens, m.linear_v.weight.detach(), m.linear_v.bias.detach())
            o = ops.gst_token_attention(q, k, v, m.h)
            return ops.gemm(o, m.linear_out.weight.detach(), m.linear_out.bias.detach())
        from ... import autograd_fns as fns

        m = self.mha
        tokens = fns.tanh(self.gst_embs)
        q = fns.linear(ref_embs, m.linear_q.weight, m.linear_q.bias, None, 1.0, None)
        k = fns.linear(tokens, m.linear_k.weight, m.linear_k.bias, None, 1.0, None)
        v = fns.linear(tokens, m.linear_v.weight, m.linear_v.bias, None, 1.0, None)
        o = fns.gst_token_attention(q, k, v, m.h)
        return fns.linear(o, m.linear_out.weight, m.linear_out.bias, None, 1.0, None)


class StyleEncoder(torch.nn.Module):
    def __init__(self, idim: int = 80, gst_tokens: int = 10, gst_token_dim: int = 256, gst_heads: int = 4,
                 conv_layers: int = 6, conv_chans_list: Sequence[int] = (32, 32, 64, 64, 128, 128),
                 conv_kernel_size: int = 3, conv_stride: int = 2, gru_layers: int = 1, gru_units: int = 128):
        super().__init__()
        self.gst_tokens = gst_tokens
        self.gst_heads = gst_heads
        self.gst_token_dim = gst_token_dim
        self.ref_enc = ReferenceEncoder(idim=idim, conv_layers=conv_layers, conv_chans_list=conv_chans_list,
                                        conv_kernel_size=conv_kernel_size, conv_stride=conv_stride,
                                        gru_layers=gru_layers, gru_units=gru_units)
        self.stl = StyleTokenLayer(ref_embed_dim=gru_units, gst_tokens=gst_tokens, gst_token_dim=gst_token_dim,
                                   gst_heads=gst_heads)

    def condition_on_gst_tokens(self, batch_size, index=0):
        if index >= self.gst_tokens:
            raise ValueError(f"We can only synthesize by conditioning on one of {self.gst_tokens} GST tokens")
        from ... import autograd_fns as fns

        mha = self.stl.mha
        g = fns.tanh_row(self.stl.gst_embs, index)  # [1, token_dim/heads]
        v = ag.linear(g, mha.linear_v.weight, mha.linear_v.bias)
        o = ag.linear(v, mha.linear_out.weight, mha.linear_out.bias)  # [1, token_dim]
        return o.expand(batch_size, -1)

    def forward(self, speech: torch.Tensor) -> torch.Tensor:
        return self.stl(self.ref_enc(speech))
