#!/usr/bin/env python
"""profiles/<tag>_ncu_full_<wl>_summary.csv and profiles/<tag>_traffic.json from the raw CSV pages of the `ncu --set
full` captures of tools/ncu_capture.sh (so that bench.py's roofline.traffic is a figure of the code it runs on).
usage: python tools/ncu_traffic.py <tag> [c2] [c3] [c5]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9,
        "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "Tbyte": 1e12}


def tag_of(name):
    """bench.py's product tags from the kernel's template arguments <NS, SIDE_T, KB, STATS, RN, PK, TFAST, MODE>."""
    m = re.search(r"project_tc_kernel<([^>]*)>", name)
    if m:
        a = [x.strip() for x in m.group(1).split(",")]
        tru = lambda x: x in ("true", "1")  # noqa: E731
        ns, side_t, stats, rn = int(a[0]), tru(a[1]), tru(a[3]), a[4]
        mode = int(a[7]) if len(a) > 7 and a[7].isdigit() else 0  # 0 fp32 field, 1 also writes the fp16 copy, 2 reads it
        t = "project_T" if side_t else ("project_S_stats" if stats else "project_S")
        if mode == 2:
            return t + "_h16"
        if mode == 1:
            return t + "_wcopy"
        return t + {1: "", 2: "_x2", 3: "_x3"}[ns] + ("_x1r" if rn in ("1", "true") else "_x1f" if rn == "2" else "")
    if "varimax_tc2_kernel" in name or "varimax_tc_kernel" in name:
        return "varimax_sweep"
    if "varimax_exact_mma_kernel" in name:
        return "varimax_sweep_fp64"
    if "gram_bf16_kernel" in name:
        return "gram_rows_bf16"
    return name.split("(")[0][-40:]


def parse(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        rec = {"kernel": d["Kernel Name"], "tag": tag_of(d["Kernel Name"])}
        for k in KEEP:
            if k in d and d[k] != "":
                try:
                    rec[k] = float(d[k].replace(",", "")) * UNIT.get(units[names.index(k)], 1.0)
                except ValueError:
                    pass
        out.append(rec)
    return out


def main():
    tag = sys.argv[1]
    traffic = {}
    for wl in sys.argv[2:] or ["c2"]:
        path = os.path.join(ROOT, "gpurun_out", f"{tag}_full_{wl}_raw.csv")
        if not os.path.exists(path):
            print("missing", path)
            continue
        recs = parse(path)
        cols = ["tag"] + KEEP + ["kernel"]
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_{wl}_summary.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols)
            for r in recs:
                w.writerow([r.get(c, "") for c in cols])
        by = {}
        for r in recs:
            if "dram__bytes_read.sum" in r:
                by.setdefault(r["tag"], []).append(r["dram__bytes_read.sum"] + r.get("dram__bytes_write.sum", 0.0))
        traffic[wl] = {t: sum(v) / len(v) for t, v in by.items()}
        for r in recs:
            print(wl, r["tag"], f"{r.get('gpu__time_duration.sum', 0) * 1e3:.3f} ms",
                  f"dram {(r.get('dram__bytes_read.sum', 0) + r.get('dram__bytes_write.sum', 0)) / 1e9:.2f} GB",
                  f"dram% {r.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f}",
                  f"tensor% {r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f}",
                  f"regs {r.get('launch__registers_per_thread', 0):.0f}")
    p = os.path.join(ROOT, "profiles", f"{tag}_traffic.json")
    old = json.load(open(p)) if os.path.exists(p) else {}
    old.update(traffic)
    old["_source"] = f"ncu --set full --clock-control none (tools/ncu_capture.sh {tag}); dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of a kind"
    json.dump(old, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
