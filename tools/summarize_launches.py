"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv ...`) into a
per-kernel share table.    python tools/summarize_launches.py gpurun_out/launches.csv "<header note>" > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        if len(r) != len(header):
            continue
        d = dict(zip(header, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((d["Kernel Name"], us))
    tot = sum(u for _, u in rows)
    agg = collections.defaultdict(lambda: [0.0, 0])
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short).replace("at::", "")
        agg[short][0] += us
        agg[short][1] += 1
    ours = sum(v[0] for k, v in agg.items() if k.startswith("k_"))
    print(f"# {note}")
    print(f"# {len(rows)} launches, total {tot / 1e3:.2f} ms; times are cold-cache and serialised under the profiler: compare SHARES, not absolutes.")
    print(f"# Kernels named k_* are this repo's (libcbops.so): {100 * ours / max(tot, 1e-9):.1f}% of the time.")
    print()
    print(f"{'share':>7} {'total_us':>10} {'launches':>8} {'avg_us':>9}  kernel")
    for k, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * us / tot:6.2f}% {us:10.1f} {cnt:8d} {us / cnt:9.1f}  {k[:150]}")


if __name__ == "__main__":
    main()
