"""Micro-benchmarks behind DESIGN.md's launch / epilogue findings: kernel-to-kernel gap and GEMM chain inside a CUDA graph
with and without programmatic dependent launch, eager launch cost, memset interaction, pre-split weights."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastspeech2_lightning_b200 import ops, _lib
x = torch.randn(1024, device="cuda")
big = torch.randn(32, 500, 256, device="cuda")
w = torch.randn(256, 256, device="cuda") / 16
def chain_small():
    y = x
    for _ in range(200):
        y = ops.tanh(y)
    return y
def chain_gemm():
    y = big
    for _ in range(50):
        y = ops.gemm(y, w, None, residual=big)
    return y
for name, fn in (("tanh x200", chain_small), ("gemm x50", chain_gemm)):
    for pdl in (1, 0, 1, 0):
        _lib.lib().fs2k_set_pdl(pdl)
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): g.replay()
        e.record(); torch.cuda.synchronize()
        print(name, "pdl", pdl, "ms per replay", s.elapsed_time(e) / 20)

# eager (no graph) launch cost with / without the PDL attribute
import time
for pdl in (1, 0, 1, 0):
    _lib.lib().fs2k_set_pdl(pdl)
    chain_small(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        chain_small()
    e.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("eager tanh x1000 pdl", pdl, "device us/kernel", s.elapsed_time(e), "host us/kernel", (t1 - t0) * 1e6 / 1000)
# a memset between kernels (what fs2k_mas_fwd / fs2k_colsum do)
for pdl in (1, 0):
    _lib.lib().fs2k_set_pdl(pdl)
    z = torch.randn(4096, 256, device="cuda")
    ops.colsum(z); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(100):
            ops.colsum(z)
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): g.replay()
    e.record(); torch.cuda.synchronize()
    print("graph memset+colsum x100 pdl", pdl, "us per pair", s.elapsed_time(e) * 1000 / 1000)

# weights pre-split in global memory vs split in shared memory at every k-block
_lib.lib().fs2k_set_pdl(1)
ws = ops.split_small(w)
def chain_gemm_pre():
    y = big
    for _ in range(50):
        y = ops.gemm(y, w, None, residual=big, w_small=ws)
    return y
ref = ops.gemm(big, w, None, residual=big)
got = ops.gemm(big, w, None, residual=big, w_small=ws)
print("presplit max |diff| vs in-kernel split:", float((ref - got).abs().max()))
for name, fn in (("gemm x50 in-kernel split", chain_gemm), ("gemm x50 presplit W", chain_gemm_pre)) * 2:
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    print(name, "us per gemm", s.elapsed_time(e) / 20 / 50 * 1000)
