#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the text summaries committed under profiles/.

    python benchmarks/ncu_summarise.py launches gpurun_out/launches_X.csv      # per-kernel shares of a launch list
    python benchmarks/ncu_summarise.py full gpurun_out/prof_X.ncu-rep          # selected `--set full` metrics per launch
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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
