"""ORACLE (test infrastructure, not product code) — the reference's acoustic-model
forward and loss restated as plain fp32 PyTorch-CPU functional code over a state dict.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference arm
may import this; the product path (`fastspeech2_lightning_b200`) never does.

Each function cites the reference `file:line` it follows (paths under /root/reference,
or torchaudio's `models/conformer.py` for the encoder/decoder, which the reference takes
from the un-vendored third-party package torchaudio — pinned 2.7.1 in `uv.lock:3632-3634`,
2.11.0 installed here; `fs2/model.py:23,95-102,112-119` are the call sites).

Parity status: PINNED.  The reference's own tests hold no golden vectors for this path
(SURVEY §4), so the pin is against outputs of the unmodified reference modules run in the
build container: `tests/golden/make_golden.py` imports `/root/reference/fs2` under a stub
shim, runs the cases, and commits inputs-by-seed + outputs under `tests/golden/`;
`tests/test_oracle.py` checks this file against those fixtures (max |Δ| ≤ 2e-5 fp32).

Everything operates on `[B, L, C]` channels-last tensors.  Dropout is not restated (parity
runs use p = 0 or eval mode); BatchNorm supports both eval and batch-stat mode.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import intops


class Cfg:
    """The few config values the arithmetic depends on (reference fs2/config/__init__.py:31-175)."""

    def __init__(self, config=None, **kw):
        m = config.model if config is not None else None
        self.heads = m.encoder.heads if m else 2
        self.layers_enc = m.encoder.layers if m else 4
        self.layers_dec = m.decoder.layers if m else 4
        self.kernel_enc = m.encoder.conv_kernel_size if m else 9
        self.kernel_dec = m.decoder.conv_kernel_size if m else 9
        self.vp_layers = m.variance_predictors.duration.n_layers if m else 5
        self.learn_alignment = m.learn_alignment if m else True
        self.use_postnet = m.use_postnet if m else True
        self.multispeaker = m.multispeaker if m else False
        self.multilingual = m.multilingual if m else False
        self.use_gst = m.use_global_style_token_module if m else False
        self.energy_level = (m.variance_predictors.energy.level.value if m else "phone")
        self.pitch_level = (m.variance_predictors.pitch.level.value if m else "phone")
        self.depthwise = m.variance_predictors.duration.depthwise if m else True
        self.pfs = bool(m and m.target_text_representation_level.value == "phonological_features")
        t = config.training if config is not None else None
        self.loss_w = dict(
            pitch=t.pitch_loss_weight if t else 0.1,
            energy=t.energy_loss_weight if t else 0.1,
            duration=t.duration_loss_weight if t else 0.1,
            mel=t.mel_loss_weight if t else 1.0,
            postnet=t.postnet_loss_weight if t else 1.0,
            ctc=t.attn_ctc_loss_weight if t else 0.1,
            bin=t.attn_bin_loss_weight if t else 0.1,
        )
        self.bin_warmup = t.attn_bin_loss_warmup_epochs if t else 100
        self.mel_loss = (m.mel_loss.value if m else "mse")
        self.var_loss = dict(
            pitch=m.variance_predictors.pitch.loss.value if m else "mse",
            energy=m.variance_predictors.energy.loss.value if m else "mse",
            duration=m.variance_predictors.duration.loss.value if m else "mse",
        )
        self.__dict__.update(kw)


# ---------------------------------------------------------------------------------------
# small pieces
# ---------------------------------------------------------------------------------------
def mask_from_lens(lens: torch.Tensor, max_len: int) -> torch.Tensor:
    """fs2/utils/heavy.py:11-15 — True on valid positions."""
    ids = torch.arange(0, int(max_len), device=lens.device, dtype=lens.dtype)
    return torch.lt(ids, lens.unsqueeze(1))


def positional_embedding(n: int, inv_freq: torch.Tensor) -> torch.Tensor:
    """fs2/layers.py:132-140 — [1, n, D] = cat(sin(p·ω), cos(p·ω))."""
    pos = torch.arange(n, dtype=torch.float32)
    s = torch.matmul(pos.unsqueeze(-1), inv_freq.unsqueeze(0))
    return torch.cat([s.sin(), s.cos()], dim=1)[None]


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def batch_norm(x_blc, sd, p, training: bool, new_stats: Optional[dict] = None):
    """nn.BatchNorm1d over channels of a [B,L,C] tensor.  Train mode: biased batch variance over all
    B·L positions *including padding* (SURVEY §8a note P); running stats: momentum 0.1, unbiased var."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    if training:
        flat = x_blc.reshape(-1, x_blc.shape[-1])
        mean = flat.mean(0)
        var = flat.var(0, unbiased=False)
        if new_stats is not None:
            n = flat.shape[0]
            new_stats[p + ".running_mean"] = 0.9 * sd[p + ".running_mean"] + 0.1 * mean.detach()
            new_stats[p + ".running_var"] = 0.9 * sd[p + ".running_var"] + 0.1 * (var.detach() * n / max(n - 1, 1))
    else:
        mean, var = sd[p + ".running_mean"], sd[p + ".running_var"]
    return (x_blc - mean) * torch.rsqrt(var + 1e-5) * w + b


def conv1d_blc(x_blc, weight, bias, padding, groups=1):
    """nn.Conv1d applied to channels-last data."""
    return F.conv1d(x_blc.transpose(1, 2), weight, bias, padding=padding, groups=groups).transpose(1, 2)


# ---------------------------------------------------------------------------------------
# Conformer — torchaudio/models/conformer.py
# ---------------------------------------------------------------------------------------
def _ffn(x, sd, p):
    """conformer.py:91-119: LN → Linear → SiLU → (Dropout) → Linear → (Dropout)."""
    h = _ln(x, sd, p + ".sequential.0")
    h = F.silu(F.linear(h, sd[p + ".sequential.1.weight"], sd[p + ".sequential.1.bias"]))
    return F.linear(h, sd[p + ".sequential.4.weight"], sd[p + ".sequential.4.bias"])


def _mhsa(x, key_pad, sd, p, heads):
    """nn.MultiheadAttention(q=k=v=x, key_padding_mask) — conformer.py:151-153,193-202.
    softmax(QKᵀ/√hd + (−inf on padded keys)) V, then out_proj.  x: [B,L,D]; key_pad True = padding."""
    B, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, L, heads, hd).transpose(1, 2)
    k = k.view(B, L, heads, hd).transpose(1, 2)
    v = v.view(B, L, heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, L, D)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _conv_module(x, sd, p, kernel, training, new_stats):
    """conformer.py:18-88: LN → pw conv (D→2D) → GLU → depthwise k (no length mask) → BatchNorm1d →
    SiLU → pw conv → (Dropout)."""
    h = _ln(x, sd, p + ".layer_norm")
    h = F.linear(h, sd[p + ".sequential.0.weight"].squeeze(-1), sd[p + ".sequential.0.bias"])
    h = F.glu(h, dim=-1)
    h = conv1d_blc(h, sd[p + ".sequential.2.weight"], sd[p + ".sequential.2.bias"], (kernel - 1) // 2, groups=h.shape[-1])
    h = batch_norm(h, sd, p + ".sequential.3", training, new_stats)
    h = F.silu(h)
    return F.linear(h, sd[p + ".sequential.5.weight"].squeeze(-1), sd[p + ".sequential.5.bias"])


def conformer(x, lengths, sd, p, n_layers, heads, kernel, training=False, new_stats=None):
    """torchaudio Conformer.forward (conformer.py:273-293) + ConformerLayer.forward (:176-212)."""
    B, L, D = x.shape
    key_pad = torch.arange(L)[None, :] >= lengths[:, None]  # :9-15
    for i in range(n_layers):
        q = f"{p}.conformer_layers.{i}"
        x = _ffn(x, sd, q + ".ffn1") * 0.5 + x  # :185-187
        x = _mhsa(_ln(x, sd, q + ".self_attn_layer_norm"), key_pad, sd, q + ".self_attn", heads) + x  # :192-203
        x = x + _conv_module(x, sd, q + ".conv_module", kernel, training, new_stats)  # :168-174
        x = _ffn(x, sd, q + ".ffn2") * 0.5 + x  # :207-209
        x = _ln(x, sd, q + ".final_layer_norm")  # :210
    return x


# ---------------------------------------------------------------------------------------
# variance adaptor pieces — fs2/variance_adaptor.py, fs2/layers.py, fs2/blocks.py
# ---------------------------------------------------------------------------------------
def variance_predictor(x, mask, sd, p, n_layers=5, depthwise=True):
    """VariancePredictor.forward variance_adaptor.py:55-62 over VarianceConvolutionLayer layers.py:20-48
    (dw conv k → 1×1 conv → ReLU → LayerNorm → (Dropout)) and DepthwiseSeparableConv1d blocks.py:4-19."""
    for i in range(n_layers):
        q = f"{p}.conv.{i}.layers"
        if depthwise:
            w0 = sd[q + ".0.module.model.0.weight"]
            k = w0.shape[-1]
            h = conv1d_blc(x, w0, sd[q + ".0.module.model.0.bias"], (k - 1) // 2, groups=x.shape[-1])
            h = F.linear(h, sd[q + ".0.module.model.1.weight"].squeeze(-1), sd[q + ".0.module.model.1.bias"])
        else:
            w0 = sd[q + ".0.module.weight"]
            h = conv1d_blc(x, w0, sd[q + ".0.module.bias"], (w0.shape[-1] - 1) // 2)
        x = _ln(F.relu(h), sd, q + ".2")
    out = F.linear(x, sd[p + ".linear.weight"], sd[p + ".linear.bias"]).squeeze(-1)
    return out * mask if mask is not None else out


def conv_attention(mel, text_emb, src_mask, prior, sd, p="variance_adaptor.attention"):
    """ConvAttention.forward attn/attention.py:195-251 as called at variance_adaptor.py:252-260:
    queries = mel [B,F,80], keys = raw text embedding [B,T,256]."""
    k = F.relu(conv1d_blc(text_emb, sd[p + ".key_proj.0.conv.weight"], sd[p + ".key_proj.0.conv.bias"], 1))
    k = conv1d_blc(k, sd[p + ".key_proj.2.conv.weight"], sd[p + ".key_proj.2.conv.bias"], 0)  # [B,T,80]
    q = F.relu(conv1d_blc(mel, sd[p + ".query_proj.0.conv.weight"], sd[p + ".query_proj.0.conv.bias"], 1))
    q = F.relu(conv1d_blc(q, sd[p + ".query_proj.2.conv.weight"], sd[p + ".query_proj.2.conv.bias"], 0))
    q = conv1d_blc(q, sd[p + ".query_proj.4.conv.weight"], sd[p + ".query_proj.4.conv.bias"], 0)  # [B,F,80]
    d = ((q[:, :, None, :] - k[:, None, :, :]) ** 2).sum(-1)  # :239  [B,F,T]
    attn = -0.0005 * d  # :241
    if prior is not None:
        attn = torch.log_softmax(attn, dim=-1) + torch.log(prior + 1e-8)  # :242-243
    attn_logprob = attn.clone()[:, None]  # :245
    # :247-248 masks .data in place: the mask is invisible to autograd but the values are -inf
    attn = attn.masked_fill((~src_mask)[:, None, :], float("-inf"))
    attn_soft = torch.softmax(attn, dim=-1)[:, None]  # :250
    return attn_soft, attn_logprob


def binarize_attention(attn_soft, in_lens, out_lens):
    """variance_adaptor.py:160-181 — log → per-item MAS on [:mel_len,:src_len] → dense 0/1."""
    with torch.no_grad():
        log_attn = torch.log(attn_soft.detach()).to(torch.float32).numpy()
        hard = intops.b_mas(log_attn, in_lens.numpy(), out_lens.numpy())
    return torch.from_numpy(hard)


def average_variance(var, durs):
    """variance_adaptor.py:207-222 — prefix-sum differences over each phone's frame span, counting only
    non-zero frames.  Floating-point (the fp32 prefix sums are order dependent), so it is
    restated with the same torch ops; `intops.average_variance` is the sequential-order variant."""
    ends = torch.cumsum(durs, dim=1).long()
    starts = F.pad(ends[:, :-1], (1, 0))
    nz = F.pad(torch.cumsum(var != 0.0, dim=1), (1, 0))
    cs = F.pad(torch.cumsum(var, dim=1), (1, 0))
    sums = (torch.gather(cs, 1, ends) - torch.gather(cs, 1, starts)).float()
    cnt = (torch.gather(nz, 1, ends) - torch.gather(nz, 1, starts)).float()
    return torch.where(cnt == 0.0, cnt, sums / cnt)


def length_regulator(x, durations, max_length):
    """variance_adaptor.py:65-81 (differentiable w.r.t. x via the gather index)."""
    _, mask, idx = intops.length_regulator(np.zeros(tuple(durations.shape) + (1,), np.float32), durations.numpy(), int(max_length))
    idx_t = torch.from_numpy(idx).long()
    valid = (idx_t >= 0)[..., None]
    out = torch.gather(x, 1, idx_t.clamp(min=0)[..., None].expand(-1, -1, x.shape[-1])) * valid
    return out, torch.from_numpy(mask)


def postnet(x, sd, training=False, new_stats=None, p="postnet"):
    """PostNet.forward layers.py:204-212: 5 × (Conv1d k5 + BatchNorm1d), tanh on the first four."""
    n = 5
    for i in range(n):
        x = conv1d_blc(x, sd[f"{p}.convolutions.{i}.0.conv.weight"], sd[f"{p}.convolutions.{i}.0.conv.bias"], 2)
        x = batch_norm(x, sd, f"{p}.convolutions.{i}.1", training, new_stats)
        if i < n - 1:
            x = torch.tanh(x)
    return x


def gst_reference_free(sd, batch_size, p="gst"):
    """StyleEncoder.condition_on_gst_tokens gst/model.py:77-85 with index 0: one key ⇒ softmax ≡ 1 ⇒
    linear_out(linear_v(tanh(gst_embs[0]))) — a constant vector."""
    g = torch.tanh(sd[p + ".stl.gst_embs"][0])
    v = F.linear(g, sd[p + ".stl.mha.linear_v.weight"], sd[p + ".stl.mha.linear_v.bias"])
    o = F.linear(v, sd[p + ".stl.mha.linear_out.weight"], sd[p + ".stl.mha.linear_out.bias"])
    return o[None].expand(batch_size, -1)


def gst_style_encoder(speech, sd, training=False, p="gst"):
    """StyleEncoder.forward gst/model.py:87-100: ReferenceEncoder (:179-199: 6×Conv2d s2 no-bias + BN2d +
    ReLU → GRU last hidden) → StyleTokenLayer (:241-257: 4-head attention over tanh(gst_embs))."""
    B = speech.shape[0]
    h = speech[:, None]
    for i in range(6):
        h = F.conv2d(h, sd[f"{p}.ref_enc.convs.{3*i}.weight"], None, stride=2, padding=1)
        q = f"{p}.ref_enc.convs.{3*i+1}"
        h = F.batch_norm(h, sd[q + ".running_mean"].clone(), sd[q + ".running_var"].clone(), sd[q + ".weight"], sd[q + ".bias"], training, 0.1, 1e-5)
        h = F.relu(h)
    h = h.transpose(1, 2).contiguous().view(B, h.shape[2], -1)
    gru = torch.nn.GRU(h.shape[-1], 128, 1, batch_first=True)
    gru.load_state_dict({k: sd[f"{p}.ref_enc.gru.{k}"] for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")})
    _, ref = torch.func.functional_call(gru, {k: sd[f"{p}.ref_enc.gru.{k}"] for k in gru.state_dict()}, (h,))
    ref = ref[-1]  # [B,128]
    # gst/attn.py:96-194 multi-head attention, query = ref (1 step), keys = values = tanh(gst_embs)
    m = p + ".stl.mha"
    tokens = torch.tanh(sd[p + ".stl.gst_embs"])  # [10,64]
    n_head, n_feat = 4, sd[m + ".linear_q.weight"].shape[0]
    dk = n_feat // n_head
    q = F.linear(ref, sd[m + ".linear_q.weight"], sd[m + ".linear_q.bias"]).view(B, n_head, 1, dk)
    k = F.linear(tokens, sd[m + ".linear_k.weight"], sd[m + ".linear_k.bias"]).view(1, -1, n_head, dk).permute(0, 2, 1, 3)
    v = F.linear(tokens, sd[m + ".linear_v.weight"], sd[m + ".linear_v.bias"]).view(1, -1, n_head, dk).permute(0, 2, 1, 3)
    a = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dk), dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, n_feat)
    return F.linear(o, sd[m + ".linear_out.weight"], sd[m + ".linear_out.bias"])


# ---------------------------------------------------------------------------------------
# full forward — fs2/model.py:153-268 with VarianceAdaptor.forward fs2/variance_adaptor.py:224-412
# ---------------------------------------------------------------------------------------
def forward(sd, cfg: Cfg, batch, stats_bins=None, inference=False, control=(1.0, 1.0, 1.0), training=False, new_stats=None,
            inject=None):
    """Returns the reference's 16-key output dict (+ `va_output`, `enc_output` for stage-wise checks).
    `control` = (pitch, energy, duration).  `training` selects BatchNorm batch statistics.
    `inject={"attn_hard": t}` replaces the MAS result (stage-wise parity: discrete decisions fed in)."""
    teacher_forcing = bool(inference and batch["mel_lens"] is not None)  # model.py:162-165
    src_lens = batch["src_lens"]
    T = int(batch["max_src_len"])
    mel_lens = batch["mel_lens"]
    max_mel_len = batch["max_mel_len"]
    src_mask = mask_from_lens(src_lens, T)
    if cfg.pfs:
        text_inputs = batch["pfs"]
        inputs = F.linear(text_inputs, sd["text_input_layer.weight"])
    else:
        text_inputs = batch["text"]
        inputs = F.embedding(text_inputs.long(), sd["text_input_layer.weight"], padding_idx=0)  # model.py:83-89,183 (pad row: no grad)
    inv_freq = sd["position_embedding.inv_freq"]
    x = inputs + positional_embedding(T, inv_freq) * src_mask.unsqueeze(2)  # :186-190
    x = conformer(x, src_lens, sd, "encoder", cfg.layers_enc, cfg.heads, cfg.kernel_enc, training, new_stats)
    enc_output = x
    if cfg.use_gst:  # :196-203
        if inference and torch.is_tensor(batch["mel_style_reference"]):
            style = gst_style_encoder(batch["mel_style_reference"], sd, training)
        elif inference and not teacher_forcing:
            style = gst_reference_free(sd, x.shape[0])
        else:
            style = gst_style_encoder(batch["mel"], sd, training)
        x = x + style.unsqueeze(1)
    if cfg.multispeaker:  # :206-208
        x = x + F.embedding(batch["speaker_id"].long(), sd["speaker_embedding.weight"]).unsqueeze(1)
    if cfg.multilingual:  # :211-213
        x = x + F.embedding(batch["language_id"].long(), sd["language_embedding.weight"]).unsqueeze(1)

    # ---- VarianceAdaptor.forward ----
    va = "variance_adaptor"
    energy_target = batch["energy"] if not inference else None
    pitch_target = batch["pitch"] if not inference else None
    dur = batch["duration"]
    duration_target = dur if (torch.is_tensor(dur) or dur[0] is not None) else None
    attn_logprob = attn_soft = attn_hard = None
    if (teacher_forcing or not inference) and cfg.learn_alignment:  # :248-305
        attn_soft, attn_logprob = conv_attention(batch["mel"], inputs, src_mask, batch["duration"], sd)
        attn_hard = binarize_attention(attn_soft, src_lens, mel_lens)
        if inject and "attn_hard" in inject:
            own_attn_hard, attn_hard = attn_hard, inject["attn_hard"]
        duration_target = attn_hard.sum(2)[:, 0, :].int()
        if energy_target is not None and cfg.energy_level == "phone":
            energy_target = average_variance(energy_target, duration_target)
        if pitch_target is not None and cfg.pitch_level == "phone":
            pitch_target = average_variance(pitch_target, duration_target)
        assert torch.all(duration_target.sum(1) == mel_lens), "BadDataError"

    def var_embed(x, target, mask, name, ctl):  # get_variance_embedding :183-205
        pred = variance_predictor(x, mask, sd, f"{va}.{name}_predictor", cfg.vp_layers, cfg.depthwise)
        bins = sd[f"{va}.{name}_bins"]
        if not inference:
            ids = torch.bucketize(target, bins)
        else:
            pred = pred * ctl
            ids = torch.bucketize(pred, bins)
        own_ids[name] = ids
        if inject and "bucket_ids" in inject:  # discrete decision fed in (stage-wise parity)
            ids = inject["bucket_ids"][name]
        return pred, F.embedding(ids, sd[f"{va}.{name}_embedding.weight"]), ids

    own_ids = {}
    energy_prediction = pitch_prediction = None
    ids_out = {}
    if cfg.energy_level == "phone":  # :309-329
        energy_prediction, e, ids_out["energy"] = var_embed(x, energy_target, src_mask, "energy", control[1])
        x = x + e
    if cfg.pitch_level == "phone":  # :330-350
        pitch_prediction, e, ids_out["pitch"] = var_embed(x, pitch_target, src_mask, "pitch", control[0])
        x = x + e
    log_dur = variance_predictor(x, src_mask, sd, f"{va}.duration_predictor", cfg.vp_layers, cfg.depthwise)  # :352
    if teacher_forcing or not inference:  # :354-358
        duration_rounded = duration_target
    else:  # :359-366
        duration_rounded = torch.clamp(torch.round(torch.exp(log_dur) - 1) * control[2], min=0).int()
    x, tgt_mask = length_regulator(x, duration_rounded, int(max_mel_len))
    if cfg.energy_level == "frame":  # :371-383
        energy_prediction, e, ids_out["energy"] = var_embed(x, energy_target, tgt_mask, "energy", control[1])
        x = x + e
    if cfg.pitch_level == "frame":  # :385-397
        pitch_prediction, e, ids_out["pitch"] = var_embed(x, pitch_target, tgt_mask, "pitch", control[0])
        x = x + e
    va_output = x

    # ---- decoder / mel / postnet — model.py:226-249 ----
    if inference and not teacher_forcing:
        mel_lens = tgt_mask.sum(1).to(torch.int32)
        max_mel_len = int(mel_lens.max())
    Fm = int(max_mel_len)
    x = x + positional_embedding(Fm, inv_freq) * tgt_mask.unsqueeze(2)
    x = conformer(x, mel_lens, sd, "decoder", cfg.layers_dec, cfg.heads, cfg.kernel_dec, training, new_stats)
    output = F.linear(x, sd["mel_linear.weight"], sd["mel_linear.bias"])
    postnet_output = output + postnet(output, sd, training, new_stats) if cfg.use_postnet else None
    return {
        "output": output,
        "postnet_output": postnet_output,
        "src_mask": src_mask,
        "src_lens": src_lens,
        "tgt_mask": tgt_mask,
        "tgt_lens": mel_lens,
        "attn_logprob": attn_logprob,
        "attn_soft": attn_soft,
        "attn_hard": attn_hard,
        "duration_prediction": log_dur,
        "duration_target": duration_target,
        "energy_prediction": energy_prediction,
        "energy_target": energy_target,
        "pitch_prediction": pitch_prediction,
        "pitch_target": pitch_target,
        "text_input": text_inputs,
        # extras for stage-wise parity
        "enc_output": enc_output,
        "va_output": va_output,
        "text_emb": inputs,
        "bucket_ids": ids_out,
        "own_bucket_ids": own_ids,
        "duration_rounded": duration_rounded,
    }


# ---------------------------------------------------------------------------------------
# losses — fs2/loss.py:19-126, fs2/attn/attention_loss.py:22-73
# ---------------------------------------------------------------------------------------
def attention_ctc_loss(attn_logprob, in_lens, out_lens, blank_logprob=-1.0):
    """attention_loss.py:29-62: prepend blank column, mask keys beyond key_len with −1e15, log_softmax,
    nn.CTCLoss(zero_infinity=True, mean) with targets 1..T."""
    T = attn_logprob.size(-1)
    lp = attn_logprob.squeeze(1).permute(1, 0, 2)
    lp = F.pad(lp, (1, 0, 0, 0, 0, 0), value=blank_logprob)
    key_inds = torch.arange(T + 1, dtype=torch.long)
    lp = lp.masked_fill(key_inds.view(1, 1, -1) > in_lens.view(1, -1, 1), -1e15)
    lp = torch.log_softmax(lp, dim=-1)
    targets = key_inds[1:].unsqueeze(0).repeat(in_lens.numel(), 1)
    return F.ctc_loss(lp, targets, input_lengths=out_lens, target_lengths=in_lens, blank=0, reduction="mean", zero_infinity=True)


def attention_bin_loss(hard, soft, eps=1e-12):
    """attention_loss.py:69-73."""
    return -torch.log(torch.clamp(soft[hard == 1], min=eps)).sum() / hard.sum()


def loss(out, batch, cfg: Cfg, current_epoch=0):
    """FastSpeech2Loss.forward loss.py:19-126 — means over the whole padded tensors."""
    fn = {"mse": F.mse_loss, "mae": F.l1_loss}
    src_mask, tgt_mask = out["src_mask"], out["tgt_mask"]
    losses = {}
    for name, level in (("pitch", cfg.pitch_level), ("energy", cfg.energy_level)):
        tgt = out[name + "_target"]
        if tgt is not None:
            m = src_mask if level == "phone" else tgt_mask
            losses[name] = fn[cfg.var_loss[name]](out[name + "_prediction"] * m, tgt * m) * cfg.loss_w[name]
    log_dt = torch.log(out["duration_target"].float() + 1) * src_mask
    losses["duration"] = fn[cfg.var_loss["duration"]](out["duration_prediction"] * src_mask, log_dt) * cfg.loss_w["duration"]
    tm = tgt_mask.unsqueeze(2)
    spec_target = batch["mel"] * tm
    losses["spec"] = fn[cfg.mel_loss](out["output"] * tm, spec_target) * cfg.loss_w["mel"]
    if cfg.use_postnet:
        losses["postnet"] = fn[cfg.mel_loss](out["postnet_output"] * tm, spec_target) * cfg.loss_w["postnet"]
    if cfg.learn_alignment:
        losses["attn_ctc"] = attention_ctc_loss(out["attn_logprob"], batch["src_lens"], batch["mel_lens"]) * cfg.loss_w["ctc"]
        w = min(current_epoch / cfg.bin_warmup, 1.0) * cfg.loss_w["bin"]
        losses["attn_bin"] = attention_bin_loss(out["attn_hard"], out["attn_soft"]) * w
    losses["total"] = sum(losses.values())
    return losses
