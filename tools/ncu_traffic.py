"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
per-kernel launches / time / DRAM bytes, and the JSON bench.py reads for `roofline.traffic` (measured, per launch).

  python tools/ncu_traffic.py profiles/r02_launches_step_b32.csv [--json profiles/r02_conv_traffic.json] [--kernel hn_conv_gemm_kernel]
"""
import argparse
import collections
import csv
import json
import re


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    col = {n: i for i, n in enumerate(header)}
    for r in rd:
        if len(r) < len(header):
            continue
        rows.append({"id": int(r[col["ID"]]), "kernel": r[col["Kernel Name"]], "metric": r[col["Metric Name"]], "unit": r[col["Metric Unit"]],
                     "value": float(r[col["Metric Value"]].replace(",", ""))})
    return rows


def to_base(v, unit):
    scale = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "s": 1.0,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1.0, "KB": 1e3, "MB": 1e6, "GB": 1e9}
    return v * scale.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json")
    ap.add_argument("--kernel", default="hn_conv_gemm_kernel")
    args = ap.parse_args()
    per = collections.defaultdict(lambda: {"t": 0.0, "rd": 0.0, "wr": 0.0})
    for r in parse(args.csv):
        d = per[(r["id"], re.sub(r"\(.*", "", r["kernel"]))]
        v = to_base(r["value"], r["unit"])
        if r["metric"].startswith("gpu__time"):
            d["t"] += v
        elif "read" in r["metric"]:
            d["rd"] += v
        elif "write" in r["metric"]:
            d["wr"] += v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, k), d in per.items():
        a = agg[k]
        a[0] += 1
        a[1] += d["t"]
        a[2] += d["rd"] + d["wr"]
    total_t = sum(a[1] for a in agg.values())
    print("%d launches, %.3f ms serialised" % (sum(a[0] for a in agg.values()), total_t * 1e3))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%8.3f ms %5d x %9.1f MB  %s" % (a[1] * 1e3, a[0], a[2] / 1e6, k[:90]))
    if args.json:
        ks = [(k, a) for k, a in agg.items() if args.kernel in k]
        n = sum(a[0] for _, a in ks)
        t = sum(a[1] for _, a in ks)
        b = sum(a[2] for _, a in ks)
        json.dump({"kernel": args.kernel, "launches": n, "dram_bytes_per_launch": b / max(n, 1), "dram_bytes_total": b, "time_ms_total": t * 1e3,
                   "share_of_step": t / total_t, "source": args.csv,
                   "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold caches, serialised)"},
                  open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
