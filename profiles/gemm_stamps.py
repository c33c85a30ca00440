"""Where the time goes inside one gemm_tc CTA: globaltimer stamps (ns) of CTA (0,0), warm caches.
slots: 0 start, 1 setup done, 2 first operands landed (MMA thread), 3 last MMA issued, 4 accumulator ready
(epilogue sees tmem_full), 5 staging done, 6 store pass done."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from fastspeech2_lightning_b200 import ops
from fastspeech2_lightning_b200._lib import lib

dev = torch.device("cuda:0")
stamps = torch.zeros(8, dtype=torch.int64, device=dev)
for mode in ("tf32", "tf32x3"):
    ops.set_precision(mode)
    for (M, K, N, res) in [(2560, 256, 256, True), (15872, 256, 256, True), (15872, 256, 1024, False), (15872, 1024, 256, True)]:
        x = torch.randn(1, M, K, device=dev); w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
        r = torch.randn(1, M, N, device=dev) if res else None
        for _ in range(3):
            ops.gemm(x, w, b, residual=r)
        lib().fs2k_gemm_tc_set_debug_stamps(stamps.data_ptr())
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); ops.gemm(x, w, b, residual=r); s1.record()
        torch.cuda.synchronize()
        lib().fs2k_gemm_tc_set_debug_stamps(None)
        t = stamps.cpu().tolist()
        rel = [round((v - t[0]) / 1000, 2) for v in t[:7]]
        print(f"{mode:7s} M={M:6d} K={K:5d} N={N:5d} res={res!s:5s} kernel={s0.elapsed_time(s1)*1000:7.1f}us stamps(us)={rel}")
