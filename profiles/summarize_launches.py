"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
per kernel: launches, total / mean duration, share of the (serialised, cold-cache) step, DRAM bytes per launch.
usage: python profiles/summarize_launches.py <launches.csv> [out.json]"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "").strip()
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per[name][row["Metric Name"]] += v
        ids[name].add(row["ID"])
    tot = sum(d["gpu__time_duration.sum"] for d in per.values())
    table = []
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[name])
        table.append({"kernel": name, "launches": n, "total_us": round(d["gpu__time_duration.sum"], 1),
                      "mean_us": round(d["gpu__time_duration.sum"] / n, 2), "share": round(d["gpu__time_duration.sum"] / tot, 4),
                      "dram_bytes_per_launch": round((d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / n)})
    res = {"source": path, "serialized_step_us": round(tot, 1), "kernels": sum(len(v) for v in ids.values()), "by_kernel": table}
    if out:
        json.dump(res, open(out, "w"), indent=1)
    print(f"serialized step {tot/1e3:.2f} ms, {res['kernels']} kernels")
    for t in table[:22]:
        print(f"{t['total_us']:9.1f} us {100*t['share']:5.1f}%  n={t['launches']:4d}  mean {t['mean_us']:8.1f} us  dram/launch {t['dram_bytes_per_launch']/1e6:8.2f} MB  {t['kernel'][:60]}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
