#!/usr/bin/env python
"""profiles/summarize_ncu.py RAW_CSV [--commit HEAD] -- per-kernel summary of an `ncu --set full ... ; ncu -i X.ncu-rep
--page raw --csv` capture: duration, DRAM bytes read + written per launch, registers, achieved occupancy, L1/L2 figures.
Writes profiles/traffic.json (what bench.py reports as roofline.traffic, keyed by its kernel-class names) and prints a
markdown table.  The unit row of the raw page (second CSV row) is honoured (Gbyte / Mbyte / byte, msecond / usecond)."""
import csv
import json
import os
import re
import sys

CLASS = [("rp_project", r"rp_project"), ("hclust", r"hclust_rnn|hclust_tri"), ("corrdist", r"corrdist_kernel"),
         ("sweep_nested", r"sweep_nested"), ("colsum", r"colsum|cellprep"), ("unit_rows", r"unit_rows")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
        "second": 1e3}
COLS = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_active"}


def main():
    path = sys.argv[1]
    commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else None
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        i = ix.get(COLS[key])
        if i is None or r[i] == "":
            return None
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    per = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        cls = next((c for c, pat in CLASS if re.search(pat, name)), None)
        if cls is None:
            continue
        d = per.setdefault(cls, {"kernel": re.sub(r"\(.*", "", name), "n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "last": None})
        d["n"] += 1
        for k in ("ms", "rd", "wr"):
            d[k] += val(r, k) or 0.0
        d["last"] = r
    out = {"source": os.path.basename(path), "commit": commit, "kernels": {}}
    print("| class | kernel | launches | ms / launch | DRAM read + write / launch | regs | warps active % | issue active % | L2 hit % | sectors / ld request |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for cls, d in per.items():
        n, r = d["n"], d["last"]
        e = {"kernel": d["kernel"], "launches": n, "ms_per_launch": d["ms"] / n, "dram_bytes_per_launch": (d["rd"] + d["wr"]) / n,
             "dram_read_per_launch": d["rd"] / n, "dram_write_per_launch": d["wr"] / n}
        for k in ("regs", "grid", "block", "warps_active_pct", "issue_active_pct", "l2_hit_pct", "dram_pct"):
            e[k] = val(r, k)
        s, q = val(r, "ld_sectors"), val(r, "ld_requests")
        e["sectors_per_ld_request"] = (s / q) if s and q else None
        out["kernels"][cls] = e
        f = lambda x, fmt="%.1f": "-" if x is None else fmt % x
        print(f"| {cls} | `{d['kernel']}` | {n} | {e['ms_per_launch']:.3f} | {e['dram_bytes_per_launch'] / 1e9:.3f} GB | {f(e['regs'], '%d')} | "
              f"{f(e['warps_active_pct'])} | {f(e['issue_active_pct'])} | {f(e['l2_hit_pct'])} | {f(e['sectors_per_ld_request'], '%.2f')} |")
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
