#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the text summaries committed under profiles/.

    python benchmarks/ncu_summarise.py launches gpurun_out/launches_X.csv      # per-kernel shares of a launch list
    python benchmarks/ncu_summarise.py full gpurun_out/prof_X.ncu-rep          # selected `--set full` metrics per launch
    python benchmarks/ncu_summarise.py traffic gpurun_out/prof_X.ncu-rep KEY KERNEL_SUBSTR rows=.. dim=.. batch=.. k=..
        # DRAM bytes per launch of the kernels whose name contains KERNEL_SUBSTR (the capture = the launches of ONE
        # step) -> entry KEY of profiles/ncu_traffic.json, which bench.py reads for roofline.traffic
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def launches(path: str) -> None:
    rows = [ln for ln in open(path, errors="replace") if ln.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3,
                  "second": 1e3}[unit]
        name = r["Kernel Name"].split("(")[0][:100]
        tot[name] += ms
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{'total ms':>10} {'launches':>8} {'share':>7}  kernel")
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{ms:10.3f} {cnt[name]:8d} {100 * ms / total:6.1f}%  {name}")
    ours = sum(ms for n, ms in tot.items() if "cmw::" in n or n.startswith(("gemm_topk", "scan_kernel", "pool_", "rescore",
                                                                            "select_", "prep_", "ingest", "multivector",
                                                                            "merge_", "exchange_", "tombstone")))
    print(f"\nkernels of libcmwdense.so: {100 * ours / total:.1f}% of all device time in the process "
          f"({total:.1f} ms over {sum(cnt.values())} launches)")


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    idx = {}
    for m in METRICS:
        for i, h in enumerate(hdr):
            if h == m or h.endswith("." + m):
                idx[m] = i
                break
    kn = hdr.index("Kernel Name")
    print("Kernel Name | " + " | ".join(m for m in METRICS if m in idx))
    print(" | " + " | ".join(units[idx[m]] for m in METRICS if m in idx))
    for r in rd[2:]:
        print(r[kn][:34] + " | " + " | ".join(r[idx[m]] for m in METRICS if m in idx))


def _raw(path: str):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    return hdr, units, rd[2:]


def _col(hdr, name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name):
            return i
    raise KeyError(name)


_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(path: str, key: str, substr: str, *kv: str) -> None:
    import json
    import os

    hdr, units, rows = _raw(path)
    kn = hdr.index("Kernel Name")
    cr, cw = _col(hdr, "dram__bytes_read.sum"), _col(hdr, "dram__bytes_write.sum")
    ct = _col(hdr, "gpu__time_duration.sum")
    extra = {}
    for name, label in (("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct_per_launch"),
                        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct_per_launch"),
                        ("lts__t_sector_hit_rate.pct", "l2_hit_rate_pct_per_launch")):
        try:
            extra[label] = _col(hdr, name)
        except KeyError:
            pass
    sel = [r for r in rows if substr in r[kn]]
    if not sel:
        raise SystemExit(f"no launch of a kernel containing {substr!r} in {path}")
    num = lambda s: float(s.replace(",", ""))  # noqa: E731
    rd_b = [num(r[cr]) * _BYTES[units[cr]] for r in sel]
    wr_b = [num(r[cw]) * _BYTES[units[cw]] for r in sel]
    meta = {}
    for item in kv:
        a, b = item.split("=", 1)
        meta[a] = int(b) if b.lstrip("-").isdigit() else b
    entry = {"kernel": sel[0][kn].split("(")[0], **meta, "launches_per_step": len(sel),
             "dram_read_bytes_per_launch": rd_b, "dram_write_bytes_per_launch": wr_b,
             "duration_per_launch": [f"{r[ct]} {units[ct]}" for r in sel],
             "traffic_bytes_per_step": sum(rd_b) + sum(wr_b),
             "source": f"{os.path.basename(path)} (`ncu --set full`), written by benchmarks/ncu_summarise.py traffic"}
    if "algorithmic_hbm_bytes_per_step" not in entry and {"rows", "dim"} <= set(meta):
        # one pass over the tiles + the row multipliers + the query tile
        elt = int(meta.get("elt_bytes", 2))  # bytes per element of the tiles the filter streams (4 for fp32 rows)
        entry["algorithmic_hbm_bytes_per_step"] = meta["rows"] * (meta["dim"] * elt + 4) + meta.get("batch", 0) * meta["dim"] * elt
    for label, col in extra.items():
        entry[label] = [num(r[col]) for r in sel]
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    with open(out_path) as f:
        doc = json.load(f)
    doc[key] = entry
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)
    print(f"{key}: {len(sel)} launches, {entry['traffic_bytes_per_step'] / 1e9:.3f} GB DRAM traffic per step "
          f"(algorithmic {entry.get('algorithmic_hbm_bytes_per_step', 0) / 1e9:.3f} GB)")


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
