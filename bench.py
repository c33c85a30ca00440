#!/usr/bin/env python
"""Benchmark of the FastSpeech2 acoustic-model hot path on B200 (contract: see DESIGN.md §measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload synth_c1|synth_c4|mas_c2|mas_c5]
    python bench.py --impl reference ...     # the reference's CPU arithmetic (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic batch.
`value` is device-timed with the inputs resident in HBM; `e2e` goes through the public module API
(`FastSpeech2.predict_step`) with pinned-host inputs copied in and the mel copied out inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[0]: base config, teacher-forced synthesis, B=16, T≈80 → F≈500
    "synth_c1": dict(kind="synth", batch=16, src=(60, 80), learn_alignment=False, metric="mel_frames_per_sec", unit="mel frames/s"),
    # configs[3]: 256 utterances of 20-200 phonemes (up to ~1500 frames), multispeaker + GST (400-frame style reference mel
    # through the GST reference encoder), dealt to the ranks as length-sorted batches of 32 (parallel.shard_utterances);
    # one step = one batch of 32, the timed steps cycle through the rank's batches
    "synth_c4": dict(kind="synth", batch=32, src=(20, 200), n_utts=256, learn_alignment=False, metric="mel_frames_per_sec", unit="mel frames/s",
                     model_kw=dict(multispeaker=True, use_global_style_token_module=True), n_speakers=8, style_frames=400,
                     desc="multispeaker + GST (400-frame style mel → reference encoder), 256 utterances of 20-200 phonemes in length-sorted batches of 32"),
    # configs[4]: long-utterance stress — one utterance of 1000 phonemes / ~8000 frames through the learned-alignment forward
    # (aligner + MAS over the 8000 × 1000 matrix + both Conformers + PostNet), eval mode, no gradient
    "synth_c5": dict(kind="synth", batch=1, src=(1000, 1000), learn_alignment=True, align_forward=True, dur_range=(6, 10),
                     metric="mel_frames_per_sec", unit="mel frames/s",
                     desc="base config, learned-alignment forward (aligner + MAS 8000x1000), one utterance of 1000 phonemes"),
    # configs[1] (N=1, fp32) / configs[2] (N>1): training step with learned alignment, batch 32 per GPU
    "train_c2": dict(kind="train", batch=32, src=(60, 80), metric="train_utts_per_sec", unit="utterances/s"),
    "mas_c2": dict(kind="mas", batch=32, F=500, T=80, metric="mas_ms_per_batch", unit="ms/batch"),
    "mas_c5": dict(kind="mas", batch=1, F=8000, T=1000, metric="mas_ms_per_batch", unit="ms/batch"),
}


DTYPES = {"fp32": "f32", "tf32": "tf32", "tf32x3": "f32 (3xTF32 tensor cores)",
          "bf16": "bf16 (tcgen05 kind::f16 operands; fp32 accumulate, residual stream, norms, softmax, master weights)",
          "bf16x3": "f32 (3xBF16 split accumulation on kind::f16 tensor cores)"}


def flops_fwd(B, T, F, aligner=False):
    """forward FLOPs on the padded B×L rectangle."""
    f = 4 * T * (3_019_264 + 1024 * T) + 4 * F * (3_019_264 + 1024 * F) + 3 * T * 663_552 + F * 40_960 + F * 8_683_520
    if aligner:
        f += T * 868_352 + F * 115_200 + F * T * 240
    return B * f


def workload_string(wl_name, wl, B, T, F):
    """One description per workload, shared by both arms (the driver compares the strings)."""
    if wl["kind"] == "train":
        return (f"{wl_name}: base config random init, training step with learned alignment (aligner+MAS, duration/pitch/energy/mel/"
                f"postnet/CTC/bin losses, backward, clip 1.0, AdamW+Noam), B={B}/GPU, T<={T}, F<={F}")
    if wl["kind"] == "synth":
        return f"{wl_name}: {wl.get('desc', 'base config random init')}, teacher-forced synthesis forward, B={B}/GPU, T<={T}, F<={F}, 80-bin mel"
    return wl_name


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_traffic(workload: str, entry_point: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel behind `entry_point`, from the
    committed ncu pass of the same workload (profiles/r1_summary_<workload>_v4.json, made by profiles/summarize_launches.py
    from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`); None when there is none."""
    names = {"fs2k_gemm_tc": ("gemm_tc_kernel",), "fs2k_mas_fwd": ("mas_dp_quad_kernel", "mas_dp_kernel"), "fs2k_attention_f32": ("attention_simt_kernel",),
             "fs2k_gemm_bf16": ("gemm_bf16_panel_kernel", "gemm_bf16_kernel"), "fs2k_attention_bf16": ("attention_tc_fwd_kernel",)}
    if entry_point not in names:
        return None
    from fastspeech2_lightning_b200 import ops as _o

    tag = "bf16" if _o.PRECISION == "bf16" else "fp32"
    for f in (ROOT / "profiles" / f"r2_summary_{workload}_{tag}.json", ROOT / "profiles" / f"r1_summary_{workload}_v4.json"):
        if not f.exists():
            continue
        ks = json.loads(f.read_text())["by_kernel"]
        hits = [k for k in ks if any(k["kernel"].endswith(n) for n in names[entry_point])]
        if hits:  # launch-weighted mean over the kernels behind the entry point
            n = sum(k["launches"] for k in hits)
            return sum(k["dram_bytes_per_launch"] * k["launches"] for k in hits) / n
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def build_model(wl, device):
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.config import FastSpeech2Config
    from fastspeech2_lightning_b200.fs2.model import FastSpeech2

    cfg = FastSpeech2Config(model=dict(learn_alignment=wl["learn_alignment"], **wl.get("model_kw", {})))
    torch.manual_seed(1234)
    model = FastSpeech2(cfg, stats=synthetic.DEFAULT_STATS,  # random init of the reference architecture
                        speaker2id={f"s{i}": i for i in range(wl.get("n_speakers", 0))})
    model.eval()
    if device is not None:
        model = model.to(device)
        if model.variance_adaptor is not None:
            model.variance_adaptor.validate_durations = False
    return cfg, model


def make_batches(wl, n, rank, world=1):
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.parallel import shard_utterances

    if wl.get("n_utts"):  # a corpus of n_utts utterances, length-sorted batches dealt round-robin to the ranks
        g = torch.Generator().manual_seed(1234)
        lengths = torch.randint(wl["src"][0], wl["src"][1] + 1, (wl["n_utts"],), generator=g).tolist()
        mine = shard_utterances(lengths, rank, world, wl["batch"])
        return [synthetic.make_batch(len(idx), wl["src"], seed=1234 + 17 * k + 1000 * rank, learn_alignment=False, inference=True,
                                     teacher_forced=True, n_speakers=wl.get("n_speakers", 1), src_lens=[lengths[i] for i in idx],
                                     style_frames=wl.get("style_frames", 0)) for k, idx in enumerate(mine)]
    out = []
    for i in range(n):
        if wl.get("align_forward"):
            b = synthetic.make_batch(wl["batch"], wl["src"], seed=1234 + 1000 * rank + i, learn_alignment=True, dur_range=wl.get("dur_range", (3, 9)))
        else:
            b = synthetic.make_batch(wl["batch"], wl["src"], seed=1234 + 1000 * rank + i, learn_alignment=wl["learn_alignment"],
                                     inference=True, teacher_forced=True)
        out.append(b)
    return out


def pin(batch):
    return {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}


def batch_bytes(batch):
    return sum(v.numel() * v.element_size() for v in batch.values() if torch.is_tensor(v))


def kernel_work(name, args):
    """(flops, bytes) of one C-ABI call, from its arguments (DESIGN.md lists the formulas)."""
    if name == "fs2k_gemm_f32" or name == "fs2k_gemm_tc":
        B, L, K, N, taps = args[2], args[3], args[4], args[6], args[7]
        return 2.0 * B * L * K * N * taps, 4.0 * (B * L * (K + N) + N * K * taps)
    if name == "fs2k_gemm_bf16":
        a16, B, L, K, N, taps = args[1], args[3], args[4], args[5], args[9], args[10]
        out_b = (4 if args[20] else 0) + (2 if args[22] else 0) + (4 if args[24] else 0) + (2 if args[25] else 0) + (4 if args[17] else 0)
        return 2.0 * B * L * K * N * taps, B * L * (K * (2 if a16 else 4) + N * out_b) + 2.0 * N * K * taps
    if name == "fs2k_gemm_wgrad_bf16":
        B, L, N, K, taps = args[4], args[5], args[6], args[7], args[8]
        return 2.0 * B * L * K * N * taps, 4.0 * (B * L * (K + N) + N * K * taps)
    if name == "fs2k_gemm_wgrad_bf16_ex":
        g16, x16, B, L, N, K, taps = args[1], args[4], args[6], args[7], args[8], args[9], args[10]
        return 2.0 * B * L * K * N * taps, B * L * (N * (2 if g16 else 4) + K * (2 if x16 else 4)) + 4.0 * N * K * taps
    if name == "fs2k_gemm_bf16_dact":
        M, K, N = args[2], args[3], args[6]
        return 2.0 * M * K * N, M * (4.0 * K + 2.0 * N + (4 if args[12] else 0) * N + (2 if args[13] else 0) * N) + 2.0 * N * K
    if name == "fs2k_attention_bf16":
        B, L, H, hd = args[2], args[3], args[4], args[5]
        return 4.0 * B * H * L * L * hd, B * L * H * hd * (3 * 2 + 4)
    if name == "fs2k_attention_bwd_bf16":
        B, L, H, hd = args[6], args[7], args[8], args[9]
        return 10.0 * B * H * L * L * hd, B * L * H * hd * (3 * 2 + 4 + 4 + 2 + 12)  # S, dP, dV, dK, dQ (the two kernels recompute S and dP)
    if name == "fs2k_attention_f32":
        B, L, H, hd = args[2], args[3], args[4], args[5]
        return 4.0 * B * H * L * L * hd, 4.0 * B * L * 4 * H * hd
    if name == "fs2k_attention_bwd_f32":
        B, L, H, hd = args[5], args[6], args[7], args[8]
        return 14.0 * B * H * L * L * hd, 4.0 * B * L * 11 * H * hd  # 7 L×L×hd products (S twice, dP twice, dV, dK, dQ)
    if name == "fs2k_gemm_wgrad" or name == "fs2k_gemm_wgrad_tc":
        B, L, N, K, taps = args[4], args[5], args[6], args[7], args[8]
        return 2.0 * B * L * K * N * taps, 4.0 * (B * L * (K + N) + N * K * taps)
    if name == "fs2k_mas_fwd":
        B, F, T = args[4], args[5], args[6]
        return 0.0, 8.0 * B * F * T
    if name == "fs2k_lr_gather":
        B, T, D, F = args[3], args[4], args[5], args[6]
        return 0.0, 4.0 * B * D * (T + F * (2 if args[8] else 1)) + 4.0 * B * T
    # memory-bound elementwise / normalisation kernels: bytes each element is read and written once
    if name == "fs2k_layernorm_fwd":
        M, D = args[4], args[5]
        return 0.0, 4.0 * M * (2 * D + 2)
    if name == "fs2k_layernorm_bwd":
        M, D = args[5], args[6]
        return 0.0, 4.0 * M * (3 * D + 2)
    if name == "fs2k_layernorm_bwd_add":
        M, D = args[5], args[6]
        return 0.0, 4.0 * M * (4 * D + 2)
    if name == "fs2k_dwconv_fwd":
        B, L, C, glu = args[2], args[3], args[4], args[8]
        return 0.0, 4.0 * B * L * C * ((2 if glu else 1) + 1)
    if name == "fs2k_dwconv_bwd":
        B, L, C, glu = args[3], args[4], args[5], args[8]
        return 0.0, 4.0 * B * L * C * (1 + 2 * (2 if glu else 1))
    if name in ("fs2k_colsum", "fs2k_colsum_bf16"):
        M, C = args[1], args[2]
        return 0.0, (2.0 if name.endswith("bf16") else 4.0) * M * C
    if name == "fs2k_act_bwd":
        M, C = args[5], args[6]
        return 0.0, 4.0 * M * C * 3
    if name in ("fs2k_affine_act", "fs2k_affine_act_bf16"):
        M, C = args[5], args[6]
        return 0.0, M * C * (4.0 + (4.0 if args[4] else 0.0) + (2.0 if name.endswith("bf16") else 4.0))
    if name in ("fs2k_bn_act_bwd", "fs2k_bn_act_bwd_bf16"):
        M, C = args[8], args[9]
        return 0.0, M * C * (2 * 8.0 + (2.0 if name.endswith("bf16") else 4.0))  # g and z are read by the statistics pass and by the apply pass
    if name == "fs2k_colstats":
        return 0.0, 4.0 * args[1] * args[2]
    if name == "fs2k_aligner_fwd":
        B, F, T, C = args[4], args[5], args[6], args[7]
        return 2.0 * B * F * T * C, 4.0 * B * (F * C + T * C + 3 * F * T)
    if name in ("fs2k_adamw_step", "fs2k_adamw_step_dev"):
        return 0.0, args[4] * (4.0 * 7 + (2.0 if args[-2] else 0.0))
    if name == "fs2k_sumsq":
        return 0.0, 4.0 * args[1]
    if name in ("fs2k_ctc_forward_sum_fwd", "fs2k_ctc_forward_sum_bwd"):
        B, F, T = (args[3], args[4], args[5]) if name.endswith("fwd") else (args[7], args[8], args[9])
        return 0.0, B * F * (4.0 * T + 8.0 * (2 * T + 1)) + (4.0 * B * F * T if name.endswith("bwd") else 0.0)
    return 0.0, 0.0


def kernel_table(by, pk, top=24):
    """Per-entry-point roofline figures of the profiled step: device time, share, achieved TFLOP/s against the measured dense
    bf16 peak and achieved algorithmic GB/s against the measured HBM peak (whichever the entry point has a formula for),
    and which of the two ceilings is the lower one for its arithmetic intensity."""
    tot = sum(d[0] for d in by.values()) or 1.0
    rows = []
    for name, d in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        ms, fl, by_, n = d
        row = {"entry": name, "launches": n, "ms": round(ms, 4), "share": round(ms / tot, 4)}
        if fl > 0:
            row["tflops"] = round(fl / (ms * 1e-3) / 1e12, 2)
            row["tensor_frac"] = round(row["tflops"] / pk["tf_sustained"], 4)
        if by_ > 0:
            row["gbs"] = round(by_ / (ms * 1e-3) / 1e9, 1)
            row["hbm_frac"] = round(row["gbs"] / pk["hbm"], 4)
        if fl > 0 and by_ > 0:
            row["lower_ceiling"] = "hbm" if by_ / (pk["hbm"] * 1e9) > fl / (pk["tf_sustained"] * 1e12) else "tensor"
        rows.append(row)
    return rows


def run_ours(args, wl_name, wl, rank, world, device):
    from fastspeech2_lightning_b200 import _lib, ops, synthetic

    pk = peaks()
    n_distinct = 4
    if wl["kind"] == "mas":
        return run_mas(args, wl_name, wl, rank, world, device, pk)
    if wl["kind"] == "train":
        return run_train(args, wl_name, wl, rank, world, device, pk)
    cfg, model = build_model(wl, device)
    host_batches = [pin(b) for b in make_batches(wl, n_distinct, rank, world)]
    n_distinct = len(host_batches)
    dev_batches = [synthetic.batch_to(b, device) for b in host_batches]
    frames = [int(b["mel_lens"].sum()) for b in host_batches]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2
    align_fwd = bool(wl.get("align_forward"))  # learned-alignment forward (mel given; aligner + MAS run), no gradient

    def step_eager(i):
        with torch.no_grad():
            return model(dev_batches[i % n_distinct]) if align_fwd else model(dev_batches[i % n_distinct], inference=True)

    l0 = ops.launch_count
    step_eager(0)
    launches_per_step = ops.launch_count - l0
    if not args.eager and not align_fwd:
        model.enable_cuda_graphs()  # public switch: predict_step replays one CUDA graph per (B,T,F) shape

    def step(i):
        # inputs already resident in HBM; with graphs: device→device copy into the captured buffers + 1 graph launch
        if align_fwd:
            return step_eager(i)
        return model.predict_step(dev_batches[i % n_distinct], i)

    from fastspeech2_lightning_b200.fs2.batching import trim_predictions

    def step_e2e(i):
        # the calls a user makes: pinned-host batch in (predict_step), every utterance's valid frames back on the host as
        # [n_mels, T] (trim_predictions = the prediction-writing callback's per-item slicing, one launch + one D2H copy)
        if align_fwd:
            with torch.no_grad():
                out = model(synthetic.batch_to(host_batches[i % n_distinct], device, non_blocking=True))
            return trim_predictions(out, model.output_key, reuse_buffer=True)
        out = model.predict_step(host_batches[i % n_distinct] if not args.eager else
                                 synthetic.batch_to(host_batches[i % n_distinct], device, non_blocking=True), i)
        return trim_predictions(out, model.output_key, reuse_buffer=True)

    for i in range(max(args.warmup, n_distinct)):
        step(i)
        step_e2e(i)
    torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        for i in range(steps):
            flush.fill_(i & 1)  # evict L2 between timed iterations (outside the event pair)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(i)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        return sum(s.elapsed_time(e) for s, e in evs)

    with ClockSampler(torch.cuda.current_device()) as clk:
        ms_dev = timed(step, args.steps)
        ms_e2e = timed(step_e2e, args.steps)
    launches = launches_per_step * args.steps  # kernels executed (graph replays run the captured launches)
    total_frames = sum(frames[i % n_distinct] for i in range(args.steps))

    # roofline pass: the same steps with every C-ABI call bracketed by CUDA events on the launch stream
    # (a spin kernel keeps the stream busy while the host queues the eager launches, so the event pairs
    # measure device time only, not host launch gaps)
    torch.cuda._sleep(int(60e-3 * 1.9e9))  # ≈ 60 ms of a spin kernel ahead of the queue (torch's own; not part of libfs2k)
    _lib.start_profile()
    for i in range(2):
        step_eager(i)
    recs = _lib.stop_profile()
    by = {}
    for name, a, ms in recs:
        f, b = kernel_work(name, a)
        d = by.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += ms; d[1] += f; d[2] += b; d[3] += 1
    tot_ms = sum(d[0] for d in by.values())
    dom = max(by, key=lambda k: by[k][0])
    d = by[dom]
    if d[1] > 0:
        roof = {"kernel": dom, "bound": "tensor", "achieved": d[1] / (d[0] * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s"}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d[2] / (d[0] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=ncu_traffic(wl_name, dom), peak_source=pk["source"], share_of_step=d[0] / tot_ms,
                launches=d[3], algorithmic_bytes_per_launch=d[2] / d[3],
                shares={k: round(v[0] / tot_ms, 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])},
                by_kernel=kernel_table(by, pk))

    # max over ranks, whole-job aggregate
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=device)
    n = torch.tensor([float(total_frames)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(n, op=torch.distributed.ReduceOp.SUM)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    all_frames = float(n[0])
    B, T, F = wl["batch"], int(host_batches[0]["max_src_len"]), int(host_batches[0]["max_mel_len"])
    line = {
        "metric": wl["metric"], "value": all_frames / (ms_dev * 1e-3), "unit": wl["unit"], "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[ops.PRECISION], "data": "synthetic",
        "config": {"workload": workload_string(wl_name, wl, B, T, F),
                   "l2": "flushed between timed iterations (256 MiB write)", "batch_per_gpu": B, "parallelism": f"replicas x{world}, no collectives", "launch": "eager" if args.eager else "cuda graph replay"},
        "e2e": {"value": all_frames / (ms_e2e * 1e-3), "unit": wl["unit"], "h2d_bytes_per_step": batch_bytes(host_batches[0]),
                "d2h_bytes_per_step": frames[0] * 80 * 4 + B * 8, "ms_per_step": ms_e2e / args.steps,
                "api": "FastSpeech2.predict_step + fs2.batching.trim_predictions (valid frames of every utterance as [n_mels, T] on the host)"},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": roof,
        "step_flops": flops_fwd(B, T, F, aligner=align_fwd), "step_tflops": flops_fwd(B, T, F, aligner=align_fwd) / (ms_dev / args.steps * 1e-3) / 1e12,
        "gpu_busy_ms_per_step": tot_ms / 2,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = reference_cpu_baseline(wl) or cpu_baseline(wl, cfg, host_batches[0])
    return line


def build_train_model(device, seed=1234):
    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.config import FastSpeech2Config
    from fastspeech2_lightning_b200.fs2.model import FastSpeech2

    cfg = FastSpeech2Config()  # base config: learned alignment, dropout 0.2 / 0.5, AdamW + Noam
    torch.manual_seed(seed)
    model = FastSpeech2(cfg, stats=synthetic.DEFAULT_STATS)
    model.train()
    if device is not None:
        model = model.to(device)
        model.variance_adaptor.validate_durations = False
        model.fused_grad_clip = 1.0  # gradient_clip_val of fs2/cli/train.py:38, fused into the optimizer launch
    return cfg, model


def make_train_batches(wl, n, rank):
    from fastspeech2_lightning_b200 import synthetic

    return [synthetic.make_batch(wl["batch"], wl["src"], seed=4321 + 1000 * rank + i, learn_alignment=True) for i in range(n)]


def run_train(args, wl_name, wl, rank, world, device, pk):
    """BASELINE configs[1]/[2]: one optimisation step = forward (aligner + MAS + both Conformers + PostNet) + losses
    + backward + gradient clip + AdamW (+ NCCL all-reduce of the flat gradient when world > 1)."""
    from fastspeech2_lightning_b200 import _lib, ops, synthetic

    cfg, model = build_train_model(device)
    (opt,), (sched,) = model.configure_optimizers()
    n_distinct = 4
    host_batches = [pin(b) for b in make_train_batches(wl, n_distinct, rank)]
    dev_batches = [synthetic.batch_to(b, device) for b in host_batches]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def step_eager(i):
        batch = dev_batches[i % n_distinct]
        opt.zero_grad()
        out = model(batch)
        losses = model.loss(out, batch, model.current_epoch)
        losses["total"].backward()
        opt.step()
        sched["scheduler"].step()
        return losses["total"]

    graphs = not args.eager

    def step(i):
        # inputs resident in HBM; with graphs: device→device copies into the captured buffers + one graph launch
        return model.optimization_step(dev_batches[i % n_distinct], use_cuda_graph=graphs)["total"]

    def step_e2e(i):
        # the call a user makes per batch: pinned-host batch in, the eight losses back on the host (Lightning's log_dict)
        losses = model.optimization_step(host_batches[i % n_distinct], use_cuda_graph=graphs)
        return torch.stack([v.detach() for v in losses.values()]).cpu()

    l0 = ops.launch_count
    step_eager(0)
    launches_per_step = ops.launch_count - l0
    for i in range(max(args.warmup, 2 * n_distinct)):  # every batch shape: first sight eager, second sight captured
        step(i)
        step_e2e(i)
    torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        for i in range(steps):
            flush.fill_(i & 1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(i)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        return sum(s.elapsed_time(e) for s, e in evs)

    with ClockSampler(torch.cuda.current_device()) as clk:
        ms_dev = timed(step, args.steps)
        ms_e2e = timed(step_e2e, args.steps)
    ms_noex = None
    if world > 1 and graphs and getattr(model, "_train_runner", None) is not None:
        # the same captured graphs replayed WITHOUT the all-reduces: what the gradient exchange costs per step at this N
        # (after the timed regions: these steps apply 1/world of the local gradient, so they come last)
        model._train_runner.skip_exchange = True
        ms_noex = timed(step, args.steps)
        model._train_runner.skip_exchange = False

    torch.cuda._sleep(int(150e-3 * 1.9e9))
    _lib.start_profile()
    step_eager(0)
    recs = _lib.stop_profile()
    by = {}
    for name, a, ms in recs:
        f, b = kernel_work(name, a)
        d = by.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += ms; d[1] += f; d[2] += b; d[3] += 1
    tot_ms = sum(d[0] for d in by.values())
    dom = max(by, key=lambda k: by[k][0])
    d = by[dom]
    if d[1] > 0:
        roof = {"kernel": dom, "bound": "tensor", "achieved": d[1] / (d[0] * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s"}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d[2] / (d[0] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=ncu_traffic(wl_name, dom), peak_source=pk["source"], share_of_step=d[0] / tot_ms,
                launches=d[3], algorithmic_bytes_per_launch=d[2] / d[3],
                shares={k: round(v[0] / tot_ms, 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]},
                by_kernel=kernel_table(by, pk))

    t = torch.tensor([ms_dev, ms_e2e, ms_noex or 0.0], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_noex = float(t[0]), float(t[1]), float(t[2])
    B, T, F = wl["batch"], int(host_batches[0]["max_src_len"]), int(host_batches[0]["max_mel_len"])
    utts = B * args.steps * world
    step_flops = 3 * flops_fwd(B, T, F, aligner=True)
    line = {
        "metric": wl["metric"], "value": utts / (ms_dev * 1e-3), "unit": wl["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPES[ops.PRECISION],
        "data": "synthetic",
        "config": {"workload": workload_string(wl_name, wl, B, T, F),
                   "data_check": "BadDataError check (sum of MAS durations == mel_lens, a host sync per step in the reference) off in the timed steps; the first eager sight of every shape runs it",
                   "l2": "flushed between timed iterations (256 MiB write)", "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": (f"dp{world}: per-rank replicas; flat gradient all-reduced over NCCL in two buckets ({'bf16' if ops.PRECISION == 'bf16' else 'fp32'} on the wire), "
                                   "the decoder/PostNet bucket overlapped with the rest of the backward") if world > 1 else "dp1",
                   "launch": ("cuda graph replay of the whole step (zero_grad, forward, losses, backward, clip, AdamW), one graph per batch shape"
                              if world == 1 else
                              "three cuda graph replays per step (zero_grad+forward+losses+backward to the decoder input | rest of the backward | clip+AdamW); "
                              "NCCL all-reduces launched between them on a communication stream")
                             if graphs else "eager (one launch per kernel)"},
        "e2e": {"value": utts / (ms_e2e * 1e-3), "unit": wl["unit"], "h2d_bytes_per_step": batch_bytes(host_batches[0]),
                "d2h_bytes_per_step": 8 * 4, "ms_per_step": ms_e2e / args.steps, "api": "FastSpeech2.optimization_step (training_step + backward + clip + FusedAdamW.step + NoamLR.step)"},
        "gpu_launches": launches_per_step * args.steps, "clocks": clk.summary(), "roofline": roof,
        "step_flops": step_flops, "step_tflops": step_flops / (ms_dev / args.steps * 1e-3) / 1e12,
        "gpu_busy_ms_per_step": tot_ms,
    }
    if world > 1 and ms_noex:
        line["exchange"] = {"ms_per_step_without_all_reduce": ms_noex / args.steps, "ms_per_step": ms_dev / args.steps,
                            "exposed_ms": (ms_dev - ms_noex) / args.steps,
                            "note": "same captured graphs replayed with the NCCL calls skipped, max over ranks: the part of the gradient exchange that is not hidden behind the backward"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = reference_cpu_baseline(wl)
        if line["cpu_baseline"] is None:
            v, secs, cores = cpu_train_baseline(wl, steps=1)
            line["cpu_baseline"] = {"value": v, "unit": wl["unit"], "cores": cores, "kind": "port",
                                    "sample": f"oracle port (torch-CPU restatement, no dropout RNG) of one full training step on one batch of the workload: {secs:.1f} s"}
        # context: the same unmodified reference modules as stock PyTorch eager on THIS GPU
        try:
            runner = reference_runner(wl, device=str(device))
            if runner is not None:
                v, secs, steps, _ = time_reference(runner, wl, 0, 6, 2, budget_s=30.0)
                line.setdefault("also", {})["reference_torch_eager_on_this_gpu"] = {
                    "metric": wl["metric"], "value": v, "unit": wl["unit"], "ms_per_step": secs * 1e3, "steps": steps,
                    "note": "UNMODIFIED reference modules moved to the B200 (stock PyTorch eager: cuBLAS / cuDNN / ATen, MAS on the host via numba as the reference does); context, not an arm of this repo"}
                del runner
                torch.cuda.empty_cache()
        except Exception as e:  # context only: never fail the bench for it
            line.setdefault("also", {})["reference_torch_eager_on_this_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return line


def run_train_stream(args, wl, device, n_steps=768, warm=512, pad_multiple=(8, 32), policies=("exact_graphs", "exact_eager", "bucketed_graphs")):
    """Shape-diverse training stream (every batch draws its own phone count and durations, as a real epoch does;
    dataset.py:257-293 pads each batch to its own maxima).  Three policies over the SAME stream, cold caches:
      exact_graphs    — the default: exact shapes, first sight eager, second sight captured, then replayed (LRU of 64 shapes);
      exact_eager     — exact shapes, eager launches only;
      bucketed_graphs — opt-in `pad_multiple` padding (fs2.batching.pad_batch_to_multiple): a few (T, F) shapes, so the
                        stream soon runs on replays.  Not reference padding: results differ as with one longer utterance.
    Reports wall-clock ms per step over the last `n_steps - warm` steps (device idle at both ends); a capture (graph
    instantiation of ≈ 850 nodes) costs ≈ 0.4 s of host time, so the count of captures inside the timed steps is reported."""
    import time

    import numpy as np

    from fastspeech2_lightning_b200 import synthetic
    from fastspeech2_lightning_b200.fs2.batching import pad_batch_to_multiple

    g = np.random.default_rng(2024)
    his = g.integers(64, 97, size=n_steps)
    stream = [synthetic.make_batch(wl["batch"], (max(int(h) - 30, 20), int(h)), seed=9000 + i, learn_alignment=True) for i, h in enumerate(his)]
    shapes = {(int(b["max_src_len"]), int(b["max_mel_len"])) for b in stream}
    out = {"steps": n_steps, "timed_steps": n_steps - warm, "distinct_shapes": len(shapes), "batch": wl["batch"], "pad_multiple": list(pad_multiple)}
    for policy in policies:
        cfg, model = build_train_model(device)
        model.train_graph_cache = 64
        model.configure_optimizers()
        batches = stream if policy != "bucketed_graphs" else [pad_batch_to_multiple(b, pad_multiple) for b in stream]
        host = [pin(b) for b in batches]
        graphs = policy != "exact_eager"
        t0, cap0 = None, 0
        for i, b in enumerate(host):
            if i == warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                cap0 = len(model._train_runner._cache)
            model.optimization_step(b, use_cuda_graph=graphs)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / (n_steps - warm)
        runner = model._train_runner
        out[policy] = {"ms_per_step": ms, "utts_per_sec": wl["batch"] / (ms * 1e-3), "captured_shapes": len(runner._cache),
                       "captures_inside_the_timed_steps": len(runner._cache) - cap0,
                       "shapes_in_stream": len({(int(b["max_src_len"]), int(b["max_mel_len"])) for b in batches})}
        del model, runner, host
        torch.cuda.empty_cache()
    return out


def cpu_train_baseline(wl, steps=1):
    """The reference's CPU training step (oracle port): forward + losses + backward + clip_grad_norm_(1.0) + AdamW."""
    from oracle import fs2_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, model = build_train_model(None)
    ocfg = fs2_oracle.Cfg(cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var", "_bins", "inv_freq")):
            v.requires_grad_(True)
            params.append(v)
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6)
    batches = make_train_batches(wl, 2, 0)

    def one(i):
        opt.zero_grad()
        out = fs2_oracle.forward(sd, ocfg, batches[i % 2], training=True, new_stats={})
        losses = fs2_oracle.loss(out, batches[i % 2], ocfg, 0)
        losses["total"].backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()

    one(0)  # warm-up
    t = sum(_time(lambda: one(i + 1)) for i in range(steps))
    return wl["batch"] * steps / t, t / steps, cores


def run_mas(args, wl_name, wl, rank, world, device, pk):
    from fastspeech2_lightning_b200 import ops

    B, F, T = wl["batch"], wl["F"], wl["T"]
    g = torch.Generator().manual_seed(1234 + rank)
    soft = torch.softmax(torch.randn(B, 1, F, T, generator=g) * 2, dim=-1)
    il = torch.full((B,), T, dtype=torch.int32)
    ol = torch.full((B,), F, dtype=torch.int32)
    hs, hil, hol = soft.pin_memory(), il.pin_memory(), ol.pin_memory()
    ds, dil, dol = soft.to(device), il.to(device), ol.to(device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def step(i):
        return ops.mas(ds, dil, dol, take_log=True, dense=True)

    def step_e2e(i):
        _, dur, hard = ops.mas(hs.to(device, non_blocking=True), hil.to(device, non_blocking=True), hol.to(device, non_blocking=True), take_log=True, dense=True)
        return dur.cpu()

    for i in range(args.warmup):
        step(i); step_e2e(i)

    def timed(fn):
        evs = []
        torch.cuda.synchronize()
        for i in range(args.steps):
            flush.fill_(i & 1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(i); e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs)

    l0 = ops.launch_count
    with ClockSampler(torch.cuda.current_device()) as clk:
        ms = timed(step)
        launches = ops.launch_count - l0
        ms_e2e = timed(step_e2e)
    per = ms / args.steps
    ach = 8.0 * B * F * T / (per * 1e-3) / 1e9
    line = {"metric": wl["metric"], "value": per, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl_name}: batched MAS (log + DP + backtrack + dense map) on [B={B},1,F={F},T={T}]", "l2": "flushed between timed iterations"},
            "e2e": {"value": ms_e2e / args.steps, "unit": wl["unit"], "h2d_bytes_per_step": B * F * T * 4 + 8 * B, "d2h_bytes_per_step": B * T * 4},
            "gpu_launches": launches, "clocks": clk.summary(),
            "roofline": {"kernel": "fs2k_mas_fwd", "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": ncu_traffic(wl_name, "fs2k_mas_fwd"), "peak_source": pk["source"]}}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import intops

        x = torch.log(soft).numpy()
        t0 = time.perf_counter(); intops.b_mas(x, il.numpy(), ol.numpy(), threads=1); t1 = time.perf_counter() - t0
        reps = max(1, int(3.0 / max(t1, 1e-3)))
        best = min(_time(lambda: intops.b_mas(x, il.numpy(), ol.numpy(), threads=1)) for _ in range(min(reps, 20)))
        line["cpu_baseline"] = {"value": best * 1e3, "unit": wl["unit"], "cores": 1, "kind": "port", "sample": f"oracle C restatement of mas_width1, serial per-item loop as in binarize_attention, same [B={B},F={F},T={T}] batch, best of {min(reps, 20)}"}
    return line


def _time(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def cpu_forward_fn(wl, cfg, batch):
    from oracle import fs2_oracle

    _, model = build_model(wl, None)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ocfg = fs2_oracle.Cfg(cfg)

    def fn():
        with torch.no_grad():
            return fs2_oracle.forward(sd, ocfg, batch, inference=True)

    return fn


def cpu_baseline(wl, cfg, batch):
    """The oracle port of the reference's CPU path on the host cores, same batch (bounded sample)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn = cpu_forward_fn(wl, cfg, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
    fn()
    best = min(_time(fn) for _ in range(3))
    frames = int(batch["mel_lens"].sum())
    return {"value": frames / best, "unit": wl["unit"], "cores": cores, "kind": "port",
            "sample": f"oracle (torch-CPU restatement of the reference forward) on one batch of the workload ({frames} frames), best of 3, {best*1e3:.0f} ms"}


def reference_runner(wl, device="cpu"):
    """The UNMODIFIED reference model (oracle/ref_runner.py over /root/reference or its build-time copy oracle/_ref) with the
    same random-init weights as our arm; None when the sources are not there (then the oracle port stands in)."""
    from oracle import ref_runner

    if not ref_runner.available():
        return None
    if wl["kind"] == "train":
        cfg, model = build_train_model(None)
    else:
        cfg, model = build_model(wl, None)
    extra = wl.get("model_kw", {})
    return ref_runner.ReferenceRunner(cfg, model.state_dict(), device=device,
                                      speaker2id={f"s{i}": i for i in range(extra.get("n_speakers", 0))},
                                      lang2id={f"l{i}": i for i in range(extra.get("n_languages", 0))})


def time_reference(runner, wl, rank, steps, warmup, budget_s=240.0):
    """(value, seconds per step, steps actually timed, description).  Steps are cut (never below 2) only if the run would
    not end within `budget_s`: a CPU training step of the real reference takes seconds."""
    if wl["kind"] == "train":
        batches = make_train_batches(wl, 2, rank)
        fn, units = runner.train_step, [wl["batch"]] * 2
    else:
        batches = make_batches(wl, 2, rank)[:2]
        if len(batches) == 1:
            batches = batches * 2
        fn = runner.align_forward if wl.get("align_forward") else runner.synthesize
        units = [int(b["mel_lens"].sum()) for b in batches]
    if runner.device.type != "cpu":
        from fastspeech2_lightning_b200 import synthetic

        batches = [synthetic.batch_to(b, runner.device) for b in batches]
    sync = (lambda: torch.cuda.synchronize()) if runner.device.type != "cpu" else (lambda: None)
    t0 = time.perf_counter()
    fn(batches[0])  # first call: numba JIT of the MAS kernels, allocator warm-up
    sync()
    for i in range(max(warmup - 1, 0)):
        fn(batches[(i + 1) % 2])
    sync()
    t = time.perf_counter()
    fn(batches[0])
    sync()
    one = time.perf_counter() - t
    steps = max(2, min(steps, int(budget_s / max(one, 1e-6))))
    t = time.perf_counter()
    done = 0
    for i in range(steps):
        fn(batches[i % 2])
        done += units[i % 2]
    sync()
    secs = time.perf_counter() - t
    return done / secs, secs / steps, steps, time.perf_counter() - t0


def run_reference(args, wl_name, wl, rank, world):
    """--impl reference: the reference's own implementation of the path on the host cores (all threads): the UNMODIFIED
    reference modules when their sources are available (kind "reference"), else the oracle port (kind "port")."""
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if wl["kind"] == "mas":
        from oracle import intops

        B, F, T = wl["batch"], wl["F"], wl["T"]
        g = torch.Generator().manual_seed(1234)
        x = torch.log_softmax(torch.randn(B, 1, F, T, generator=g) * 2, dim=-1).numpy()
        il, ol = [T] * B, [F] * B
        for _ in range(args.warmup):
            intops.b_mas(x, il, ol)
        t = sum(_time(lambda: intops.b_mas(x, il, ol)) for _ in range(args.steps))
        v = t / args.steps * 1e3
        return {"impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": v, "higher_is_better": False, "dtype": "f32", "data": "synthetic", "config": {"workload": wl_name},
                "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": cores, "kind": "port", "sample": "oracle b_mas (OpenMP over items)"},
                "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    b0 = (make_train_batches if wl["kind"] == "train" else make_batches)(wl, 1, 0)[0]
    B, T, F = wl["batch"], int(b0["max_src_len"]), int(b0["max_mel_len"])
    runner = reference_runner(wl)
    if runner is not None:
        v, secs, steps, total = time_reference(runner, wl, 0, args.steps, args.warmup)
        kind = "reference"
        sample = (f"{steps} timed steps (of {args.steps} asked; {args.warmup} warm-up) of the same workload through the UNMODIFIED reference modules "
                  f"(fs2.model.FastSpeech2 under oracle/ref_shim.py, torch CPU, numba MAS, dropout active in training), {secs:.2f} s/step, {total:.0f} s in all")
    elif wl["kind"] == "train":
        steps = min(args.steps, 2)
        v, secs, cores = cpu_train_baseline(wl, steps=steps)
        kind, sample = "port", f"{steps} training steps of the same workload on the host CPU (oracle port; dropout RNG not included; reference sources not found)"
    else:
        cfg, _ = build_model(wl, None)
        batches = make_batches(wl, 2, 0)
        fns = [cpu_forward_fn(wl, cfg, b) for b in batches]
        steps = min(args.steps, 8)
        for i in range(min(args.warmup, 2)):
            fns[i % 2]()
        t = sum(_time(fns[i % 2]) for i in range(steps))
        v = sum(int(batches[i % 2]["mel_lens"].sum()) for i in range(steps)) / t
        secs = t / steps
        kind, sample = "port", f"{steps} steps of the same workload on the host CPU (oracle port; reference sources not found)"
    return {"impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"], "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(wl_name, wl, B, T, F)},
            "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def reference_cpu_baseline(wl, budget_s=25.0):
    """cpu_baseline of the main arm: a bounded sample (≈ 10–30 s of CPU work) of the same workload through the real reference."""
    runner = reference_runner(wl)
    if runner is None:
        return None
    v, secs, steps, total = time_reference(runner, wl, 0, 4, 1, budget_s=budget_s)
    return {"value": v, "unit": wl["unit"], "cores": os.cpu_count() or 1, "kind": "reference",
            "sample": f"{steps} steps of the same workload through the UNMODIFIED reference modules on the host CPU (dropout active in training), "
                      f"{secs:.2f} s/step after a JIT / warm-up step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train_c2", choices=sorted(WORKLOADS))
    ap.add_argument("--also", default="train_c2@bf16,train_stream@bf16,synth_c1,synth_c1@bf16,synth_c4@bf16,synth_c5@bf16,mas_c2,mas_c5", help="extra workloads measured briefly and attached under 'also' (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs: one Python-driven launch per kernel")
    ap.add_argument("--precision", default=None, choices=["fp32", "tf32", "tf32x3", "bf16", "bf16x3"],
                    help="default: tf32x3 (fp32-parity arithmetic, BASELINE configs[1]) on one GPU, bf16 (configs[2]) on several")
    ap.add_argument("--backward-precision", default=None, choices=["fp32", "tf32", "tf32x3", "bf16", "bf16x3"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        line = run_reference(args, args.workload, wl, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a collective that some rank never joins aborts the job after two minutes instead of holding the box until the driver's limit
        torch.distributed.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=120))
    from fastspeech2_lightning_b200 import ops as _ops

    if args.precision is None:
        # BASELINE.json: configs[1] = the fp32 training step on one B200; configs[2] = bf16 training, batch-sharded DDP at 2/4/8
        args.precision = "bf16" if (world > 1 and wl["kind"] == "train") else "tf32x3"
    _ops.set_precision(args.precision, args.backward_precision)
    line = run_ours(args, args.workload, wl, rank, world, device)
    if world == 1 and args.also:
        # the other two headline numbers of BASELINE.json's metric (mel frames/s synthesized, MAS ms/batch)
        import copy

        line.setdefault("also", {})
        for entry in [n for n in args.also.split(",") if n and n != args.workload]:
            # "synth_c1@tf32": the same workload in single-pass TF32; "synth_c1@dec-tf32": only the decoder / PostNet
            # (everything after the last discrete decision) in single-pass TF32 — the reduced-precision configuration
            name, _, prec = entry.partition("@")
            a2 = copy.copy(args)
            a2.no_cpu_baseline = True
            a2.steps = min(args.steps, 10)
            if prec.startswith("dec-"):
                _ops.set_precision(args.precision, args.backward_precision, decoder=prec[4:])
            else:
                _ops.set_precision(prec or args.precision, args.backward_precision)
            try:
                if name == "train_stream":  # shape-diverse stream: exact shapes vs opt-in bucketed padding (run_train_stream)
                    line["also"][entry] = run_train_stream(a2, WORKLOADS["train_c2"], device)
                    line["also"][entry]["dtype"] = DTYPES[_ops.PRECISION]
                    continue
                sub = run_ours(a2, name, WORKLOADS[name], rank, world, device)
                if prec:
                    sub["dtype"] = f"{sub.get('dtype')} / decoder+postnet {prec[4:]} (reduced-precision mode, mel L1 <= 1e-2)" if prec.startswith("dec-") else sub.get("dtype")
                line["also"][entry] = {k: sub[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "config", "dtype", "gpu_launches") if k in sub}
            except Exception as e:  # an extra line must never cost the headline
                line["also"][entry] = {"failed": f"{type(e).__name__}: {e}"[:300]}
            finally:
                _ops.set_precision(args.precision, args.backward_precision)
                torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators that were captured into CUDA graphs do not tear down reliably (destroy_process_group
        # was seen to hang after the result was printed): drain the device, meet the other ranks, leave at once.
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
