"""Per-shape timing of the contraction kernels inside a CUDA graph (N launches back to back, L2 state as in a real
step: the previous launch's output is the next one's residual stream).  Prints µs per launch, achieved TFLOP/s and
the HBM floor of the shape.  usage: python profiles/gemm_shapes.py [bf16|tf32x3] [n_launches]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from fastspeech2_lightning_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda"
HBM = 6541.8e9
M = 32 * 500


def bench(name, fn, flops, bytes_):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    print(f"{name:46s} {us:8.2f} us   {flops / us / 1e6:8.1f} TFLOP/s   HBM floor {bytes_ / HBM * 1e6:6.2f} us ({bytes_ / 1e6:6.1f} MB)", flush=True)


def lin(K, N, act=None, residual=False, dropout=0.0, silu_pair=False, Mrows=M):
    x = torch.randn(Mrows, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    res = torch.randn(Mrows, N, device=dev) if residual else None
    if mode == "bf16":
        w16, _ = ops.cast_bf16(w)
        if silu_pair:
            return lambda: ops.gemm_bf16(x, w16, b, act="silu", want_pre="fp32", dropout_p=dropout, seed=1)
        return lambda: ops.gemm_bf16(x, w16, b, act=act, residual=res, dropout_p=dropout, seed=1)
    ops.set_precision(mode)
    return lambda: ops.gemm(x, w, b, act=act, residual=res, dropout_p=dropout, seed=1)


def dgrad(Nf, Kf, Mrows=M):
    g = torch.randn(Mrows, Nf, device=dev)
    w = torch.randn(1, Nf, Kf, device=dev) / Nf ** 0.5
    if mode == "bf16":
        w16, _ = ops.cast_bf16(w)
        return lambda: ops.gemm_bf16(g, w16, None, w_mn=True)
    ops.set_precision(mode)
    wt = ops.weight_taps_transposed(w)
    return lambda: ops.gemm(g, wt, None)


def wgrad(N, K, taps=1, B=32, L=500):
    g = torch.randn(B, L, N, device=dev)
    x = torch.randn(B, L, K, device=dev)
    acc = torch.zeros((N, K, taps) if taps > 1 else (N, K), device=dev)
    ops.set_precision(mode)
    return lambda: ops.gemm_wgrad(g, x, taps, (taps - 1) // 2, taps > 1, accumulate_into=acc)


def conv(K, N, taps, B=32, L=500, a16=False, hint=None):
    x = torch.randn(B, L, K, device=dev)
    w = torch.randn(taps, N, K, device=dev) / (K * taps) ** 0.5
    if mode == "bf16":
        w16, _ = ops.cast_bf16(w)
        src = ops.cast_bf16(x)[0] if a16 else x
        h = hint if hint is not None else (256 if N % 256 == 0 else 0)
        return lambda: ops.gemm_bf16(src, w16, None, taps_pad=(taps - 1) // 2, block_n_hint=h)
    ops.set_precision(mode)
    return lambda: ops.gemm(x, w, None, taps_pad=(taps - 1) // 2)


F4 = 4.0
print(f"mode {mode}, M = {M}")
bench("linear 256->256 +bias+residual", lin(256, 256, residual=True), 2.0 * M * 256 * 256, M * 256 * F4 * 3)
bench("linear 256->256 +bias+residual+dropout", lin(256, 256, residual=True, dropout=0.1), 2.0 * M * 256 * 256, M * 256 * F4 * 3)
bench("linear 256->768 (qkv)", lin(256, 768), 2.0 * M * 256 * 768, M * (256 + 768) * F4)
bench("linear 256->512 (pw1)", lin(256, 512), 2.0 * M * 256 * 512, M * (256 + 512) * F4)
bench("linear 256->1024 silu pair", lin(256, 1024, silu_pair=True), 2.0 * M * 256 * 1024, M * (256 + 2048) * F4)
bench("linear 256->1024 silu pair + dropout", lin(256, 1024, silu_pair=True, dropout=0.1), 2.0 * M * 256 * 1024, M * (256 + 2048) * F4)
bench("linear 1024->256 +residual", lin(1024, 256, residual=True), 2.0 * M * 256 * 1024, M * (1024 + 512) * F4)
bench("linear 1024->256 +residual+dropout", lin(1024, 256, residual=True, dropout=0.1), 2.0 * M * 256 * 1024, M * (1024 + 512) * F4)
bench("dgrad g[256] . W[256,1024] -> 1024", dgrad(256, 1024), 2.0 * M * 256 * 1024, M * (256 + 1024) * F4)
bench("dgrad g[1024] . W[1024,256] -> 256", dgrad(1024, 256), 2.0 * M * 256 * 1024, M * (256 + 1024) * F4)
bench("dgrad g[256] . W[256,256] -> 256", dgrad(256, 256), 2.0 * M * 256 * 256, M * 512 * F4)
bench("encoder linear 256->1024 (M=2560)", lin(256, 1024, silu_pair=True, Mrows=2560), 2.0 * 2560 * 256 * 1024, 2560 * (256 + 2048) * F4)
bench("encoder linear 1024->256 (M=2560)", lin(1024, 256, residual=True, Mrows=2560), 2.0 * 2560 * 256 * 1024, 2560 * (1024 + 512) * F4)
bench("postnet conv5 512->512", conv(512, 512, 5), 2.0 * M * 512 * 512 * 5, M * 1024 * F4)
bench("postnet conv5 80->512", conv(80, 512, 5), 2.0 * M * 80 * 512 * 5, M * 592 * F4)
if mode == "bf16":
    bench("postnet conv5 512->512, bf16 A by TMA, BN 256", conv(512, 512, 5, a16=True), 2.0 * M * 512 * 512 * 5, M * (512 * 2 + 512 * 4))
    bench("postnet conv5 512->512, bf16 A by TMA, BN 128", conv(512, 512, 5, a16=True, hint=-1), 2.0 * M * 512 * 512 * 5, M * (512 * 2 + 512 * 4))
    xa = ops.cast_bf16(torch.randn(M, 1024, device=dev))[0]
    wa = ops.cast_bf16(torch.randn(256, 1024, device=dev) / 32)[0]
    ra = torch.randn(M, 256, device=dev)
    bench("linear 1024->256 +residual, bf16 A by TMA", lambda: ops.gemm_bf16(xa, wa, None, residual=ra), 2.0 * M * 256 * 1024, M * (1024 * 2 + 512 * 4))
    xb = torch.randn(M, 256, device=dev)
    wb = ops.cast_bf16(torch.randn(1024, 256, device=dev) / 16)[0]
    bench("linear 256->1024 silu, bf16 outputs (pre + value)", lambda: ops.gemm_bf16(xb, wb, None, act="silu", want_c=False, want_c16=True, want_pre="bf16"),
          2.0 * M * 256 * 1024, M * (256 * 4 + 2048 * 2))
    bench("linear 256->1024 silu pair, tile-per-CTA kernel", lambda: ops.gemm_bf16(xb, wb, None, act="silu", want_pre="fp32", block_n_hint=-1),
          2.0 * M * 256 * 1024, M * (256 + 2048) * F4)
bench("wgrad N=1024 K=256", wgrad(1024, 256), 2.0 * M * 256 * 1024, M * 1280 * F4)
bench("wgrad N=256 K=1024", wgrad(256, 1024), 2.0 * M * 256 * 1024, M * 1280 * F4)
bench("wgrad N=256 K=256", wgrad(256, 256), 2.0 * M * 256 * 256, M * 512 * F4)
bench("wgrad conv5 N=512 K=512", wgrad(512, 512, 5), 2.0 * M * 512 * 512 * 5, M * 1024 * F4)
