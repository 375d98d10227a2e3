#!/usr/bin/env python
"""launch_table.py — per-kernel table of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`): launches, summed device time, share of the frame, DRAM bytes per launch.

  python tools/launch_table.py gpurun_out/r02_cfg3_launches.csv [--frame-end k_resolve:N]  > profiles/r02_cfg3_launches.md

Only the launches of the FIRST frame are counted: the list is cut after the N-th `k_resolve` (N = passes per frame).
Times under ncu are cold-cache and serialised: compare the SHARES with bench.py's `kernels`, not the absolutes.
"""
import argparse
import collections
import csv
import json
import re
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--passes", type=int, default=0, help="k_resolve launches that end the first frame (0 = all launches)")
    ap.add_argument("--json", default=None, help="also write {kernel: {launches, ms, dram_bytes_per_launch}} here")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.csv, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    seen_resolve = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("void ", "").replace("rt::", "").strip()
        metric, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        d = per.setdefault(name, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
        d["ids"].add(r[col["ID"]])
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if metric == "gpu__time_duration.sum":
            d["ns"] += val * scale
            if name.startswith("k_resolve"):
                seen_resolve += 1
        elif metric == "dram__bytes_read.sum":
            d["rd"] += val * scale
        elif metric == "dram__bytes_write.sum":
            d["wr"] += val * scale
        if a.passes and seen_resolve >= a.passes and metric == "dram__bytes_write.sum" and name.startswith("k_resolve"):
            break
    tot = sum(d["ns"] for d in per.values()) or 1.0
    print(f"# launch list `{a.csv}`" + (f" (first frame: up to the {a.passes}th k_resolve)" if a.passes else ""))
    print("\n| kernel | launches | device ms (sum) | share | DRAM MB / launch (read + write) |\n|---|---|---|---|---|")
    out = {}
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(d["ids"])
        print(f"| `{name}` | {n} | {d['ns'] / 1e6:.2f} | {100 * d['ns'] / tot:.1f} % | {(d['rd'] + d['wr']) / n / 1e6:.2f} |")
        out[name] = {"launches": n, "ms": d["ns"] / 1e6, "dram_bytes_per_launch": (d["rd"] + d["wr"]) / n}
    print(f"\ntotal {tot / 1e6:.1f} ms over {sum(len(d['ids']) for d in per.values())} launches")
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
