#!/usr/bin/env python
"""summarise_ncu.py — turn an `ncu --set full --import-source on` report into the markdown tables kept under profiles/.

  python tools/summarise_ncu.py gpurun_out/r02b_cfg3_full.ncu-rep [--top 12] > profiles/r02_cfg3_full.md

Reads the report HERE (no GPU needed) through `ncu -i … --page raw --csv` (per-launch metrics) and `--page source --csv`
(per-SASS-instruction samples and execution counts).  Per launch it prints duration, registers, occupancy, issue and pipe
utilisation, the stall reasons above 0.15 per issue, DRAM bytes, and (for the first launch of every kernel) the `--top`
hottest SASS instructions by stall samples plus an opcode histogram weighted by executed instructions.
"""
import argparse
import collections
import csv
import io
import re
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "inst pipe_fma %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "inst pipe_alu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "inst pipe_xu (MUFU) %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "inst pipe_lsu %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "inst pipe_fp64 %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sass__inst_executed_local_loads", "local (spill) loads"),
    ("sass__inst_executed_local_stores", "local (spill) stores"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("rt::", "")
    return name.strip()


def raw_tables(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in data:
        d = {"id": r[col["ID"]], "kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        vals = collections.OrderedDict()
        for key, label in RAW_KEYS:
            if key in col:
                v = fnum(r[col[key]])
                if v is not None:
                    vals[label] = (v, units[col[key]])
        stalls = []
        for h, i in col.items():
            m = re.match(r"smsp__average_warps?_issue_stalled_([a-z_]+)_per_issue_active\.ratio", h)
            if m and m.group(1) != "selected":
                v = fnum(r[i])
                if v is not None and v >= 0.15:
                    stalls.append((v, m.group(1)))
        d["vals"], d["stalls"] = vals, sorted(stalls, reverse=True)
        launches.append(d)
    return launches


def source_table(rep, kernel_regex, top):
    rows = ncu_csv(rep, "source", ["--kernel-name", f"regex:{kernel_regex}", "--launch-count", "1"])
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hi is None:
        return None
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) - 2 or not r[0].startswith("0x") or r[0] in seen:
            continue
        seen.add(r[0])
        data.append(r)

    def I(r, k):
        try:
            return int(r[col[k]])
        except (ValueError, KeyError):
            return 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(I(r, "# Samples") for r in data) or 1
    tot_ex = sum(I(r, "Instructions Executed") for r in data) or 1
    hot = sorted(data, key=lambda r: -I(r, "# Samples"))[:top]
    lines = []
    for r in hot:
        st = sorted(((I(r, x), x[6:]) for x in stall_cols if I(r, x) > 0), reverse=True)[:3]
        lines.append((r[col["Source"]].strip(), 100.0 * I(r, "# Samples") / tot, I(r, "Instructions Executed"),
                      ", ".join(f"{n} {c * 100 // max(I(r, '# Samples'), 1)}%" for c, n in st)))
    ops, ops_s = collections.Counter(), collections.Counter()
    for r in data:
        t = r[col["Source"]].strip().split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0].rstrip(";")
        ops[op] += I(r, "Instructions Executed")
        ops_s[op] += I(r, "# Samples")
    mix = [(op, 100.0 * n / tot_ex, 100.0 * ops_s[op] / tot) for op, n in ops.most_common(14)]
    return lines, mix, len(data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--title", default=None)
    a = ap.parse_args()
    launches = raw_tables(a.report)
    print(f"# {a.title or a.report}\n")
    print("Per-launch metrics (`ncu --set full --clock-control none`, `--page raw`); durations are profiler-replay times, never bench values.\n")
    for d in launches:
        print(f"## launch {d['id']}: `{d['kernel']}`  grid {d['grid']} block {d['block']}\n")
        print("| metric | value |\n|---|---|")
        for label, (v, u) in d["vals"].items():
            print(f"| {label} | {v:,.3f} {u} |" if abs(v) < 1e6 else f"| {label} | {v:,.0f} {u} |")
        if d["stalls"]:
            print("| stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in d["stalls"]) + " |")
        print()
    done = set()
    for d in launches:
        if d["kernel"] in done:
            continue
        done.add(d["kernel"])
        res = source_table(a.report, re.escape(d["kernel"].split("<")[0]), a.top)
        if not res:
            continue
        lines, mix, n = res
        print(f"## `{d['kernel']}`: hottest SASS instructions of its first captured launch ({n} instructions; `--page source`)\n")
        print("| SASS | % of stall samples | executed (warp-level) | top stall reasons |\n|---|---|---|---|")
        for src, pct, ex, st in lines:
            print(f"| `{src}` | {pct:.2f} | {ex:,} | {st} |")
        print("\nopcode mix: " + ", ".join(f"{op} {p:.1f}% of executed / {s:.1f}% of samples" for op, p, s in mix) + "\n")


if __name__ == "__main__":
    sys.exit(main())
