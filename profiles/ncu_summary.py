"""Condenses an `ncu --set full` report into a small CSV (one column per profiled launch) with the metrics the roofline /
DESIGN numbers are read from.  usage: python profiles/ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or any(h.startswith(k) for k in KEEP)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch {r[hdr.index('ID')]}" for r in data])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i][:110] for r in data])
    print(out, len(cols), "metrics x", len(data), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
