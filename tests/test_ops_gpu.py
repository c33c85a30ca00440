"""Op-level parity of the CUDA kernels (through the C ABI) against the oracle / golden KATs.
Integer and index results must be bit-exact; fp32 results within rel 1e-4 of max|ref| (north_star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN
from oracle import fs2_oracle, intops

pytestmark = pytest.mark.gpu
KATS = dict(np.load(GOLDEN / "kats.npz", allow_pickle=False))
FP32_TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def ops():
    from fastspeech2_lightning_b200 import ops as o

    return o


def close(got, want, tol=FP32_TOL, what=""):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    want = want.detach().cpu().double().numpy() if torch.is_tensor(want) else np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got)), f"{what}: finite pattern differs"
    if not fin.any():
        return
    scale = max(float(np.abs(want[fin]).max()), 1e-6)
    err = float(np.abs(got[fin] - want[fin]).max()) / scale
    assert err <= tol, f"{what}: max err / max|ref| = {err:.3e} > {tol}"


# ------------------------------------------------------------------------------------------------
# MAS — bit exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(len([k for k in KATS if k.startswith("mas.in.")])))
def test_mas_kats(i):
    x, want = KATS[f"mas.in.{i}"], KATS[f"mas.out.{i}"]
    n_mel, n_text = x.shape
    if n_text == 1:
        pytest.skip("T=1 is undefined behaviour in the reference")
    t = torch.from_numpy(x).to(dev())[None, None]
    path, dur, hard = ops().mas(t, torch.tensor([n_text], device=dev()), torch.tensor([n_mel], device=dev()))
    assert np.array_equal(hard[0, 0].cpu().numpy(), want)
    assert np.array_equal(dur[0].cpu().numpy(), want.sum(0).astype(np.int32))
    assert np.array_equal(path[0].cpu().numpy(), want.argmax(1).astype(np.int32))


def test_mas_batched_ragged_matches_reference_b_mas():
    x = torch.from_numpy(KATS["bmas.in"]).to(dev())
    il, ol = torch.from_numpy(KATS["bmas.in_lens"]).to(dev()), torch.from_numpy(KATS["bmas.out_lens"]).to(dev())
    _, dur, hard = ops().mas(x, il, ol)
    assert np.array_equal(hard.cpu().numpy(), KATS["bmas.out"])
    assert np.array_equal(dur.sum(1).cpu().numpy(), KATS["bmas.out_lens"])


@pytest.mark.parametrize("B,F,T,seed", [(32, 500, 80, 0), (4, 1500, 200, 1), (3, 700, 1100, 2), (2, 300, 2500, 3),
                                        (1, 8000, 1000, 4), (5, 33, 32, 5), (2, 64, 1024, 6), (3, 257, 77, 7), (3, 100, 130, 8),
                                        (2, 90, 4096, 9), (4, 41, 3, 10), (2, 9, 515, 11), (3, 2, 2, 12)])
@pytest.mark.parametrize("wavefront", [1, 0])  # 1: four columns per thread, skewed blocks of 8 frames (default); 0: one barrier per frame
def test_mas_random_vs_oracle(B, F, T, seed, wavefront, request):
    from fastspeech2_lightning_b200._lib import lib

    lib().fs2k_mas_set_wavefront(wavefront)
    request.addfinalizer(lambda: lib().fs2k_mas_set_wavefront(1))
    g = np.random.default_rng(seed)
    x = g.standard_normal((B, 1, F, T)).astype(np.float32)
    # quantise a part to force ties, sprinkle -inf
    x[:, :, ::3] = np.round(x[:, :, ::3] * 2) / 2
    x[g.random(x.shape) < 0.01] = -np.inf
    il = g.integers(max(2, T // 2), T + 1, size=B).astype(np.int32)
    ol = g.integers(max(2, F // 2), F + 1, size=B).astype(np.int32)
    il[0], ol[0] = T, F
    if B > 2:
        il[-1], ol[-1] = 2, max(2, F // 3)   # a text much shorter than the padded width: whole warps right of it
    want = intops.b_mas(x, il, ol)
    path, dur, hard = ops().mas(torch.from_numpy(x).to(dev()), torch.from_numpy(il).to(dev()), torch.from_numpy(ol).to(dev()))
    assert np.array_equal(hard.cpu().numpy(), want)
    assert np.array_equal(dur.cpu().numpy(), want[:, 0].sum(1).astype(np.int32))
    p = path.cpu().numpy()
    for b in range(B):
        assert np.all(p[b, ol[b]:] == -1)
        assert np.all(np.diff(p[b, : ol[b]]) >= 0) and np.all(np.diff(p[b, : ol[b]]) <= 1)


def test_mas_fused_log_matches_separate_log():
    g = np.random.default_rng(9)
    B, F, T = 6, 210, 47
    soft = torch.softmax(torch.from_numpy(g.standard_normal((B, 1, F, T)).astype(np.float32)) * 3, dim=-1)
    il = torch.tensor([47, 40, 33, 20, 47, 9], dtype=torch.int32)
    ol = torch.tensor([210, 100, 180, 77, 150, 30], dtype=torch.int32)
    want = intops.b_mas(torch.log(soft).numpy(), il.numpy(), ol.numpy())
    _, dur, hard = ops().mas(soft.to(dev()), il.to(dev()), ol.to(dev()), take_log=True)
    got = hard.cpu().numpy()
    # CUDA logf vs torch CPU log may differ in the last ulp; alignments may only differ at exact near-ties
    assert (got != want).sum() <= 2 * 2
    assert np.array_equal(dur.sum(1).cpu().numpy(), ol.numpy())


# ------------------------------------------------------------------------------------------------
# LengthRegulator — bit-exact index / mask / copy
# ------------------------------------------------------------------------------------------------
def _lr(x, d, max_len, inv_freq=None):
    o = ops()
    cum, total = o.lr_scan(d)
    width = min(int(total.max()), int(max_len))
    return o.lr_gather(x, cum, total, width, inv_freq, want_idx=True), total


@pytest.mark.parametrize("tag", "abcd")
def test_length_regulator_kats(tag):
    x = torch.from_numpy(KATS[f"lr.{tag}.x"]).to(dev())
    if x.shape[-1] % 4:
        x = x.repeat(1, 1, 4 // x.shape[-1] if x.shape[-1] < 4 else 1)
    d = torch.from_numpy(KATS[f"lr.{tag}.d"]).to(dev())
    (out, _, mask, idx), _ = _lr(x.contiguous(), d, int(KATS[f"lr.{tag}.maxlen"]))
    want_out, want_mask = KATS[f"lr.{tag}.out"], KATS[f"lr.{tag}.mask"]
    assert np.array_equal(mask.cpu().numpy(), want_mask)
    assert np.array_equal(out.cpu().numpy()[..., : want_out.shape[-1]], want_out)


@pytest.mark.parametrize("B,T,D,maxlen,seed", [(16, 80, 256, 10**6, 0), (7, 200, 256, 900, 1), (3, 2500, 64, 10**6, 2), (1, 1000, 256, 10**6, 3)])
def test_length_regulator_random(B, T, D, maxlen, seed):
    g = np.random.default_rng(seed)
    x = g.standard_normal((B, T, D)).astype(np.float32)
    d = g.integers(0, 10, size=(B, T)).astype(np.int32)
    d[g.random((B, T)) < 0.2] = 0
    if B > 2:
        d[1, T // 2:] = 0
    want_out, want_mask, want_idx = intops.length_regulator(x, d, maxlen)
    inv_freq = (1 / (10000 ** (torch.arange(0.0, D, 2.0) / D))).to(dev())
    (out, out_pos, mask, idx), total = _lr(torch.from_numpy(x).to(dev()), torch.from_numpy(d).to(dev()), maxlen, inv_freq)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(mask.cpu().numpy(), want_mask)
    assert np.array_equal(out.cpu().numpy(), want_out)
    assert np.array_equal(total.cpu().numpy(), d.sum(1))
    pe = fs2_oracle.positional_embedding(want_out.shape[1], inv_freq.cpu())
    want_pos = torch.from_numpy(want_out) + pe * torch.from_numpy(want_idx >= 0)[..., None]
    close(out_pos, want_pos, 2e-6, "lr out_pos")


# ------------------------------------------------------------------------------------------------
# bucketize / embedding add / average_variance / rounding / embeddings
# ------------------------------------------------------------------------------------------------
def test_bucketize_kat_and_embed_add():
    o = ops()
    v, bins, want = torch.from_numpy(KATS["bucket.v"]), torch.from_numpy(KATS["bucket.bins"]), KATS["bucket.ids"]
    ids = o.bucketize(v.to(dev()), bins.to(dev()))
    assert ids.dtype == torch.int64 and np.array_equal(ids.cpu().numpy(), want)
    g = torch.Generator().manual_seed(0)
    table, x = torch.randn(256, 256, generator=g), torch.randn(v.numel(), 256, generator=g)
    y, ids2, vs = o.bucketize_embed_add(v.to(dev()), bins.to(dev()), table.to(dev()), x.to(dev()), scale=1.0, want_scaled=True)
    assert np.array_equal(ids2.cpu().numpy(), want)
    assert torch.equal(y.cpu(), x + table[torch.from_numpy(want)])
    # inference: prediction·control is bucketized and returned
    y, ids3, vs = o.bucketize_embed_add(v.to(dev()), bins.to(dev()), table.to(dev()), x.to(dev()), scale=1.3, want_scaled=True)
    assert torch.equal(vs.cpu().nan_to_num(7.0), (v * 1.3).nan_to_num(7.0))
    assert np.array_equal(ids3.cpu().numpy(), torch.bucketize(v * 1.3, bins).numpy())


@pytest.mark.parametrize("tag", "ab")
def test_average_variance(tag):
    var, dur, want = KATS[f"avg.{tag}.var"], KATS[f"avg.{tag}.dur"], KATS[f"avg.{tag}.out"]
    cum, _ = ops().lr_scan(torch.from_numpy(dur).to(dev()))
    got = ops().average_variance(torch.from_numpy(var).to(dev()), cum).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    assert np.array_equal(got == 0, want == 0)


@pytest.mark.parametrize("tag", "abc")
def test_round_durations(tag):
    x, c, want = KATS[f"round.{tag}.in"], float(KATS[f"round.{tag}.control"]), KATS[f"round.{tag}.out"]
    got = ops().round_durations(torch.from_numpy(x).to(dev()), c).cpu().numpy()
    # an exact .5 boundary depends on the last ulp of exp(): compare away from the boundaries
    frac = np.abs((np.exp(x.astype(np.float64)) - 1) % 1 - 0.5)
    ok = ~(frac < 1e-4)
    assert np.array_equal(got[ok], want[ok])
    g = np.random.default_rng(0)
    r = (g.standard_normal(4000) * 1.5).astype(np.float32)
    frac = np.abs((np.exp(r.astype(np.float64)) - 1) % 1 - 0.5)
    ok = frac > 1e-4
    assert np.array_equal(ops().round_durations(torch.from_numpy(r).to(dev()), c).cpu().numpy()[ok], intops.round_durations(r, c)[ok])


def test_embed_posenc_and_masks():
    o = ops()
    g = np.random.default_rng(1)
    B, T, D, n = 5, 37, 256, 40
    text = torch.from_numpy(g.integers(0, n, size=(B, T)).astype(np.int32))
    lens = torch.tensor([37, 20, 1, 30, 11], dtype=torch.int32)
    table = torch.from_numpy(g.standard_normal((n, D)).astype(np.float32))
    inv_freq = torch.from_numpy(KATS["posenc.inv_freq"])
    emb, x = o.embed_posenc(text.to(dev()), table.to(dev()), inv_freq.to(dev()), lens.to(dev()))
    want_emb = table[text.long()]
    mask = fs2_oracle.mask_from_lens(lens, T)
    assert torch.equal(emb.cpu(), want_emb)
    close(x, want_emb + fs2_oracle.positional_embedding(T, inv_freq) * mask[..., None], 2e-6, "embed+posenc")
    assert torch.equal(o.lens_mask(lens.to(dev()), T).cpu(), mask)
    assert torch.equal(o.mask_lens(mask.to(dev())).cpu(), lens)
    # long positions (C5: 8000 frames) keep fp32 accuracy
    z = torch.zeros(1, 8200, D)
    pe = o.add_posenc(z.to(dev()), inv_freq.to(dev()), torch.tensor([8200], dtype=torch.int32, device=dev()))
    np.testing.assert_allclose(pe.cpu().numpy()[0, ::41], KATS["posenc.out"], rtol=0, atol=2e-6)
    rows = torch.from_numpy(g.standard_normal((7, D)).astype(np.float32))
    ids = torch.tensor([3, 0, 6, 6, 1], dtype=torch.int32)
    style = torch.from_numpy(g.standard_normal((B, D)).astype(np.float32))
    y = o.add_rows(x, [(style.to(dev()), None), (rows.to(dev()), ids.to(dev()))])
    close(y, x.cpu() + style[:, None] + rows[ids.long()][:, None], 1e-6, "add_rows")


# ------------------------------------------------------------------------------------------------
# floating-point kernels vs a plain PyTorch fp32 reference of the same op
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,D", [(1, 256), (1000, 256), (77, 512), (300, 80), (5, 1024)])
def test_layernorm(M, D):
    g = torch.Generator().manual_seed(M)
    x, w, b = torch.randn(M, D, generator=g) * 3 + 1, torch.randn(D, generator=g), torch.randn(D, generator=g)
    y, mean, rstd = ops().layernorm(x.to(dev()), w.to(dev()), b.to(dev()), 1e-5, save_stats=True)
    close(y, F.layer_norm(x, (D,), w, b, 1e-5), 1e-5, "layernorm")
    close(mean, x.mean(-1), 1e-5, "ln mean")


@pytest.mark.parametrize("B,L,K,N,taps,act,extra", [
    (2, 100, 256, 1024, 1, "silu", ""), (16, 80, 1024, 256, 1, None, "res_alpha"), (3, 77, 256, 768, 1, None, ""),
    (2, 130, 80, 512, 5, "tanh", "bn"), (2, 130, 512, 80, 5, None, "bn_res"), (4, 33, 256, 512, 3, "relu", ""),
    (1, 500, 39, 256, 1, None, "nobias"), (2, 64, 160, 80, 1, "relu", ""), (2, 9, 512, 512, 5, "tanh", "bn"),
    (1, 1, 256, 80, 1, None, ""), (3, 50, 256, 256, 3, "relu", "mask"),
])
def test_gemm_conv_epilogues(B, L, K, N, taps, act, extra):
    o = ops()
    g = torch.Generator().manual_seed(B * 1000 + L)
    x = torch.randn(B, L, K, generator=g)
    w = torch.randn(N, K, taps, generator=g) / (K * taps) ** 0.5
    bias = None if "nobias" in extra else torch.randn(N, generator=g) * 0.1
    ref = F.conv1d(x.transpose(1, 2), w, bias, padding=(taps - 1) // 2).transpose(1, 2)
    kw = {}
    if "bn" in extra:
        sc, sh = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1
        ref = ref * sc + sh
        kw.update(scale=sc.to(dev()), shift=sh.to(dev()))
    if act:
        ref = {"silu": F.silu, "relu": F.relu, "tanh": torch.tanh}[act](ref)
    if "res" in extra:
        res = torch.randn(B, L, N, generator=g)
        alpha = 0.5 if "alpha" in extra else 1.0
        ref = ref * alpha + res
        kw.update(residual=res.to(dev()), alpha=alpha)
    if "mask" in extra:
        rm = torch.rand(B, L, generator=g) > 0.3
        ref = ref * rm[..., None]
        kw.update(row_mask=rm.to(dev()))
    wt = o.conv_weight_taps(w.to(dev()))
    got = o._gemm_f32(x.to(dev()), wt, None if bias is None else bias.to(dev()), taps_pad=(taps - 1) // 2, act=act, **kw)
    close(got, ref, 2e-5, f"gemm {B}x{L}x{K}->{N} taps{taps} {act} {extra}")


def test_rowdot():
    g = torch.Generator().manual_seed(3)
    x, w, b = torch.randn(4, 50, 256, generator=g), torch.randn(256, generator=g) / 16, torch.randn(1, generator=g)
    mask = torch.rand(4, 50, generator=g) > 0.2
    got = ops().rowdot(x.to(dev()), w.to(dev()), b.to(dev()), mask.to(dev()))
    close(got, (x @ w + b) * mask, 1e-5, "rowdot")


@pytest.mark.parametrize("B,L,H,hd,lens", [(3, 100, 2, 128, [100, 37, 64]), (2, 500, 2, 128, [500, 123]), (1, 65, 4, 64, [65]),
                                           (2, 31, 2, 128, [1, 31]), (1, 1100, 2, 128, [1000])])
def test_attention(B, L, H, hd, lens):
    g = torch.Generator().manual_seed(L)
    D = H * hd
    qkv = torch.randn(B, L, 3 * D, generator=g)
    lens_t = torch.tensor(lens, dtype=torch.int32)
    q, k, v = [t.view(B, L, H, hd).transpose(1, 2) for t in qkv.split(D, -1)]
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    s = s.masked_fill((torch.arange(L)[None, :] >= lens_t[:, None])[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, D)
    got, lse = ops().attention(qkv.to(dev()), lens_t.to(dev()), H, want_lse=True)
    close(got, ref, 2e-5, "attention")
    close(lse, torch.logsumexp(s, -1), 2e-5, "attention lse")


@pytest.mark.parametrize("K,glu,bn", [(9, True, True), (9, True, False), (3, False, False), (5, False, True), (31, True, True)])
def test_dwconv(K, glu, bn):
    g = torch.Generator().manual_seed(K)
    B, L, C = 3, 150, 256
    x = torch.randn(B, L, 2 * C if glu else C, generator=g)
    w, b = torch.randn(C, 1, K, generator=g) / K ** 0.5, torch.randn(C, generator=g) * 0.1
    h = F.glu(x, -1) if glu else x
    ref = F.conv1d(h.transpose(1, 2), w, b, padding=(K - 1) // 2, groups=C).transpose(1, 2)
    kw = {}
    if bn:
        sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
        ref = F.silu(ref * sc + sh)
        kw = dict(scale=sc.to(dev()), shift=sh.to(dev()))
    got = ops().dwconv(x.to(dev()), w.to(dev()), b.to(dev()), channels=C, glu=glu, **kw)
    close(got, ref, 2e-5, "dwconv")


def test_batchnorm_training_statistics_and_running_update():
    o = ops()
    g = torch.Generator().manual_seed(5)
    z = torch.randn(7, 90, 256, generator=g) * 2 + 0.5
    bn = torch.nn.BatchNorm1d(256)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(256, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(256, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(256, generator=g) * 0.1)
    ref_bn = torch.nn.BatchNorm1d(256)
    ref_bn.load_state_dict(bn.state_dict())
    ref_bn.train()
    ref = F.silu(ref_bn(z.transpose(1, 2)).transpose(1, 2))
    bn = bn.to(dev())
    scale, shift = o.bn_scale_shift(bn, z.to(dev()), True)
    got = o.affine_act(z.to(dev()), scale, shift, "silu")
    close(got, ref, 2e-5, "bn train")
    close(bn.running_mean, ref_bn.running_mean, 1e-5, "running_mean")
    close(bn.running_var, ref_bn.running_var, 1e-5, "running_var")
    assert int(bn.num_batches_tracked) == 1
    ref_bn.eval()
    scale, shift = o.bn_scale_shift(bn, None, False)
    close(o.affine_act(z.to(dev()), scale, shift, None), ref_bn(z.transpose(1, 2)).transpose(1, 2), 2e-5, "bn eval")


@pytest.mark.parametrize("B,F_,T,lens", [(3, 100, 20, [20, 11, 17]), (2, 333, 65, [65, 30]), (1, 700, 1000, [1000])])
def test_aligner_scores(B, F_, T, lens):
    g = torch.Generator().manual_seed(T)
    q, k = torch.randn(B, F_, 80, generator=g) * 3, torch.randn(B, T, 80, generator=g) * 3
    prior = torch.rand(B, F_, T, generator=g)
    lens_t = torch.tensor(lens, dtype=torch.int32)
    d = -0.0005 * ((q[:, :, None] - k[:, None]) ** 2).sum(-1)
    lp = torch.log_softmax(d, -1) + torch.log(prior + 1e-8)
    pad = torch.arange(T)[None, :] >= lens_t[:, None]
    soft_ref = torch.softmax(lp.masked_fill(pad[:, None, :], float("-inf")), -1)
    soft, logprob = ops().aligner_scores(q.to(dev()), k.to(dev()), prior.to(dev()), lens_t.to(dev()))
    close(logprob[:, 0], lp, 1e-5, "attn_logprob")
    close(soft[:, 0], soft_ref, 1e-5, "attn_soft")
    soft2, lp2 = ops().aligner_scores(q.to(dev()), k.to(dev()), None, None)
    close(lp2[:, 0], d, 1e-5, "no prior logprob")
    close(soft2[:, 0], torch.softmax(d, -1), 1e-5, "no prior soft")


def test_losses():
    o = ops()
    g = torch.Generator().manual_seed(8)
    pred, tgt = torch.randn(4, 60, 80, generator=g), torch.randn(4, 60, 80, generator=g)
    mask = torch.rand(4, 60, generator=g) > 0.3
    for kind, fn in (("mse", F.mse_loss), ("mae", F.l1_loss)):
        ref = fn(pred * mask[..., None], tgt * mask[..., None]) * 0.7
        close(o.masked_loss_fwd(pred.to(dev()), tgt.to(dev()), mask.to(dev()), kind, 0.7), ref, 1e-5, kind)
    dur = torch.randint(0, 9, (4, 60), generator=g, dtype=torch.int32)
    p1 = torch.randn(4, 60, generator=g)
    ref = F.mse_loss(p1 * mask, torch.log(dur.float() + 1) * mask) * 0.1
    close(o.masked_loss_fwd(p1.to(dev()), dur.to(dev()), mask.to(dev()), "mse", 0.1, log1p_int_target=True), ref, 1e-5, "dur")
    soft = torch.softmax(torch.randn(3, 1, 50, 12, generator=g), -1)
    hard = F.one_hot(torch.randint(0, 12, (3, 1, 50), generator=g), 12).float()
    hard[2, 0, 40:] = 0
    ref = -torch.log(torch.clamp(soft[hard == 1], min=1e-12)).sum() / hard.sum()
    loss, _ = o.bin_loss_fwd(hard.to(dev()), soft.to(dev()), 1e-12)
    close(loss, ref, 1e-5, "bin loss")


# ------------------------------------------------------------------------------------------------
# tcgen05 tensor-core GEMM (TMA + TMEM): same contract as the FFMA kernel
# ------------------------------------------------------------------------------------------------
TC_CASES = [
    (2, 100, 256, 1024, 1, "silu", ""), (16, 80, 1024, 256, 1, None, "res_alpha"), (3, 77, 256, 768, 1, None, ""),
    (2, 130, 80, 512, 5, "tanh", "bn"), (2, 130, 512, 80, 5, None, "bn_res"), (4, 33, 256, 512, 3, "relu", ""),
    (2, 64, 160, 80, 1, "relu", ""), (2, 9, 512, 512, 5, "tanh", "bn"), (1, 1, 256, 80, 1, None, ""),
    (3, 50, 256, 256, 3, "relu", "mask"), (5, 300, 256, 256, 1, None, "res_ln"), (2, 140, 1024, 256, 1, None, "res_alpha_ln2"),
    (1, 3000, 256, 256, 1, "relu", "ln"),
]


@pytest.mark.parametrize("mode,tol", [("tf32x3", 6e-5), ("tf32", 3e-3)])
@pytest.mark.parametrize("B,L,K,N,taps,act,extra", TC_CASES)
def test_gemm_tensor_core(mode, tol, B, L, K, N, taps, act, extra):
    o = ops()
    g = torch.Generator().manual_seed(B * 1000 + L + N)
    x = torch.randn(B, L, K, generator=g)
    w = torch.randn(N, K, taps, generator=g) / (K * taps) ** 0.5
    bias = torch.randn(N, generator=g) * 0.1
    ref = F.conv1d(x.transpose(1, 2), w, bias, padding=(taps - 1) // 2).transpose(1, 2)
    kw = {}
    if "bn" in extra:
        sc, sh = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1
        ref = ref * sc + sh
        kw.update(scale=sc.to(dev()), shift=sh.to(dev()))
    if act:
        ref = {"silu": F.silu, "relu": F.relu, "tanh": torch.tanh}[act](ref)
    if "res" in extra:
        res = torch.randn(B, L, N, generator=g)
        alpha = 0.5 if "alpha" in extra else 1.0
        ref = ref * alpha + res
        kw.update(residual=res.to(dev()), alpha=alpha)
    if "mask" in extra:
        rm = torch.rand(B, L, generator=g) > 0.3
        ref = ref * rm[..., None]
        kw.update(row_mask=rm.to(dev()))
    ln_ref = ln2_ref = None
    if "ln" in extra:
        g1, b1 = torch.randn(N, generator=g), torch.randn(N, generator=g)
        ln_ref = F.layer_norm(ref, (N,), g1, b1, 1e-5)
        kw.update(ln=(g1.to(dev()), b1.to(dev()), 1e-5))
        if "ln2" in extra:
            g2, b2 = torch.randn(N, generator=g), torch.randn(N, generator=g)
            ln2_ref = F.layer_norm(ln_ref, (N,), g2, b2, 1e-5)
            kw.update(ln2=(g2.to(dev()), b2.to(dev())))
    wt = o.conv_weight_taps(w.to(dev()))
    prev = o.PRECISION
    o.set_precision(mode)
    try:
        got = o.gemm(x.to(dev()), wt, bias.to(dev()), taps_pad=(taps - 1) // 2, act=act, **kw)
    finally:
        o.set_precision(prev)
    what = f"gemm_tc[{mode}] {B}x{L}x{K}->{N} taps{taps} {act} {extra}"
    if ln_ref is None:
        close(got, ref, tol, what)
    else:
        close(got[0], ref, tol, what)
        close(got[1], ln_ref, tol * 5, what + " ln")
        if ln2_ref is not None:
            close(got[2], ln2_ref, tol * 5, what + " ln2")


def test_zero_sized_calls_are_no_ops():
    """Empty batches / zero rows must return empty results without touching the GPU state (every C entry point returns
    FS2K_OK before its null checks when there is nothing to do)."""
    from fastspeech2_lightning_b200 import ops

    dev = "cuda:0"
    w = torch.randn(256, 256, device=dev)
    b = torch.randn(256, device=dev)
    assert ops.gemm(torch.empty(0, 256, device=dev), w, b).shape == (0, 256)
    assert ops.gemm(torch.empty(0, 17, 256, device=dev), w, b, residual=torch.empty(0, 17, 256, device=dev)).shape == (0, 17, 256)
    assert ops.layernorm(torch.empty(0, 256, device=dev), b, b, 1e-5).shape == (0, 256)
    qkv = torch.empty(0, 9, 768, device=dev)
    assert ops.attention(qkv, torch.empty(0, dtype=torch.int32, device=dev), 2).shape == (0, 9, 256)
    path, dur, hard = ops.mas(torch.empty(0, 1, 12, 5, device=dev), torch.empty(0, dtype=torch.int32, device=dev),
                              torch.empty(0, dtype=torch.int32, device=dev))
    assert hard.shape == (0, 1, 12, 5) and dur.shape == (0, 5)
    cum, total = ops.lr_scan(torch.empty(0, 7, dtype=torch.int32, device=dev))
    assert cum.shape == (0, 7) and total.shape == (0,)
    assert ops.colsum(torch.empty(0, 256, device=dev)).abs().sum() == 0
    e = torch.empty(0, 256, device=dev)
    dx, dg, db = ops.layernorm_bwd(e, e, torch.empty(0, device=dev), torch.empty(0, device=dev), b)
    assert dx.shape == (0, 256) and float(dg.abs().sum()) == 0 and float(db.abs().sum()) == 0
    # an utterance of zero phones / zero frames inside a batch: durations all zero → nothing expanded, mask all False
    dur = torch.tensor([[2, 1, 0], [0, 0, 0]], dtype=torch.int32, device=dev)
    cum, total = ops.lr_scan(dur)
    assert total.tolist() == [3, 0]
    torch.cuda.synchronize()


def test_gemm_with_presplit_weights_is_bit_identical():
    """3×TF32 with the weights' small parts pre-computed in global memory (fs2k_split_small, loaded by TMA) must give
    exactly what the in-kernel split gives — same operands, same MMA order."""
    from fastspeech2_lightning_b200 import ops

    g = torch.Generator().manual_seed(21)
    for (B, L, K, N) in [(3, 130, 256, 256), (2, 77, 256, 1024), (4, 64, 1024, 256), (1, 40, 64, 80)]:
        x = torch.randn(B, L, K, generator=g).to("cuda:0")
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to("cuda:0")
        b = torch.randn(N, generator=g).to("cuda:0")
        ws = ops.split_small(w)
        big = w.view(torch.int32).bitwise_and(-8192).view(torch.float32)  # 0xFFFFE000
        assert torch.equal(ws, w - big)
        a = ops.gemm(x, w, b, act="silu")
        c = ops.gemm(x, w, b, act="silu", w_small=ws)
        assert torch.equal(a, c), (B, L, K, N)


def test_attention_longest_first_dispatch_order_changes_nothing():
    """`order` only permutes which CTA works on which utterance: forward and backward must be bit-identical."""
    from fastspeech2_lightning_b200 import ops

    g = torch.Generator().manual_seed(4)
    B, L, H, hd = 7, 150, 2, 128
    qkv = (torch.randn(B, L, 3 * H * hd, generator=g) * 0.5).to("cuda:0")
    lens = torch.tensor([150, 33, 97, 150, 1, 64, 120], dtype=torch.int32, device="cuda:0")
    order = ops.attention_order(lens)
    assert order.tolist() == [0, 3, 6, 2, 5, 1, 4]
    o0, lse0 = ops.attention(qkv, lens, H, want_lse=True)
    o1, lse1 = ops.attention(qkv, lens, H, want_lse=True, order=order)
    assert torch.equal(o0, o1) and torch.equal(lse0, lse1)
    go = torch.randn(B, L, H * hd, generator=g).to("cuda:0")
    assert torch.equal(ops.attention_bwd(qkv, o0, lse0, go, lens, H), ops.attention_bwd(qkv, o0, lse0, go, lens, H, order=order))
