"""A few launches of the representative contraction kernels — the command the `ncu --set full` captures wrap.
usage: python profiles/gemm_one.py [panel|tile|conv|wgrad|attn|tf32x3]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from fastspeech2_lightning_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "panel"
M = 32 * 500
dev = "cuda"
x = torch.randn(M, 256, device=dev)
res = torch.randn(M, 256, device=dev)
if what in ("panel", "tile"):
    hint = 0 if what == "panel" else -1
    w, _ = ops.cast_bf16(torch.randn(256, 256, device=dev) / 16)
    w1, _ = ops.cast_bf16(torch.randn(1024, 256, device=dev) / 16)
    b, b1 = torch.randn(256, device=dev), torch.randn(1024, device=dev)
    for _ in range(3):
        ops.gemm_bf16(x, w, b, residual=res, block_n_hint=hint)                          # decoder out-projection shape
        ops.gemm_bf16(x, w1, b1, act="silu", want_pre="fp32", block_n_hint=hint)          # FFN first linear: value + pre-activation
elif what == "conv":
    a = torch.randn(32, 500, 512, device=dev)
    w, _ = ops.cast_bf16(torch.randn(5, 512, 512, device=dev) / 50)
    a16, _ = ops.cast_bf16(a)
    for _ in range(3):
        ops.gemm_bf16(a, w, None, taps_pad=2, block_n_hint=256)      # PostNet conv, fp32 activations
        ops.gemm_bf16(a16, w, None, taps_pad=2, block_n_hint=256)    # same, bf16 activations by TMA
elif what == "wgrad":
    g = torch.randn(32, 500, 1024, device=dev)
    xx = torch.randn(32, 500, 256, device=dev)
    for _ in range(3):
        ops.gemm_wgrad_bf16(g, xx, 1, 0, False)
elif what == "attn":
    qkv16, _ = ops.cast_bf16(torch.randn(32, 500, 768, device=dev) * 0.5)
    lens = torch.randint(400, 501, (32,), device=dev, dtype=torch.int32)
    dout = torch.randn(32, 500, 256, device=dev)
    for _ in range(3):
        out, lse = ops.attention_bf16(qkv16, lens, 2, want_lse=True)
        ops.attention_bwd_bf16(qkv16, out, lse, dout, lens, 2)
else:
    ops.set_precision("tf32x3")
    w = torch.randn(256, 256, device=dev) / 16
    b = torch.randn(256, device=dev)
    y = x
    for _ in range(6):
        y = ops.gemm(y, w, b, residual=x)
torch.cuda.synchronize()
print("ok", what)
