#!/usr/bin/env python
"""Summarise ncu CSV logs (run here, on the CPU box) into the tables committed under profiles/.

  launch list   : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv <cmd>
      python scripts/ncu_summarise.py shares L.csv [--skip N] [--steps K]
  GEMM traffic  : ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
                  --clock-control none -k regex:'gemm_kernel|mlp_chain' --csv --log-file T.csv <cmd running ONE eager step after warm-up>
      python scripts/ncu_summarise.py traffic T.csv --launches-per-step 24 > profiles/gemm_traffic.json
"""
import argparse
import collections
import csv
import json
import re
import sys


def read(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    return rows


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").strip()


def val(r):
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    return v, u


def to_us(v, u):
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6, "nsecond": v / 1e3, "usecond": v, "msecond": v * 1e3}.get(u, v / 1e3)


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["shares", "traffic"])
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0, help="launches to drop at the front (warm-up)")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--launches-per-step", type=int, default=0, help="traffic: keep only the LAST n launches")
    a = ap.parse_args()
    rows = read(a.csv)
    by_id = collections.OrderedDict()
    for r in rows:
        by_id.setdefault(r["ID"], {"name": short(r["Kernel Name"])})[r["Metric Name"]] = val(r)
    launches = list(by_id.values())[a.skip:]
    if a.mode == "shares":
        tot = collections.Counter(); cnt = collections.Counter()
        for l in launches:
            us = to_us(*l["gpu__time_duration.sum"])
            tot[l["name"]] += us; cnt[l["name"]] += 1
        total = sum(tot.values())
        print(f"{len(launches)} launches, {total / a.steps:.1f} us of kernel time per step ({a.steps} step(s)) under ncu\n")
        print("| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|")
        for k, v in tot.most_common():
            print(f"| `{k[:90]}` | {cnt[k] / a.steps:g} | {v / a.steps:.1f} | {100 * v / total:.1f}% |")
        return
    if a.launches_per_step:
        launches = launches[-a.launches_per_step:]
    out = {"launches_per_step": len(launches), "by_kernel": {}}
    tb = tu = 0.0
    agg = collections.OrderedDict()
    for l in launches:
        b = to_bytes(*l["dram__bytes_read.sum"]) + to_bytes(*l["dram__bytes_write.sum"])
        us = to_us(*l["gpu__time_duration.sum"])
        tp = [v for k, (v, _) in ((k, x) for k, x in l.items() if k.startswith("sm__pipe_tensor"))]
        d = agg.setdefault(l["name"], {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "tensor_w": 0.0})
        d["launches"] += 1; d["dram_bytes"] += b; d["us"] += us; d["tensor_w"] += (tp[0] if tp else 0.0) * us
        tb += b; tu += us
    for k, d in agg.items():
        out["by_kernel"][k] = {"launches": d["launches"], "dram_bytes": d["dram_bytes"], "us": d["us"],
                               "dram_GBs": d["dram_bytes"] / d["us"] / 1e3, "tensor_pipe_active_pct_time_weighted": d["tensor_w"] / d["us"]}
    out.update(dram_bytes_per_step=tb, dram_bytes_per_launch=tb / max(len(launches), 1), kernel_us_per_step_under_ncu=tu,
               algorithmic_flops_per_step=3733400322048)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
