import os, sys, torch
sys.path.insert(0, "/root/repo")
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
