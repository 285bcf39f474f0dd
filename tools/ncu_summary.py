"""Summarise an .ncu-rep (or the launch-list csv) into the few numbers the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep            # per-launch key metrics
    python tools/ncu_summary.py --launches gpurun_out/launches.csv       # per-kernel share of the step
"""

from __future__ import annotations

import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__cycles_active.avg", "sm__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]  # fmt: skip


def launches(path: str) -> None:
    lines = [line for line in open(path) if not line.startswith("==")]
    agg: dict[str, list] = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (ncu: cold-cache, serialised - compare shares)")
    print(f"{'share':>7} {'launches':>8} {'avg us':>10}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{100 * a[1] / tot:6.2f}% {a[0]:8d} {a[1] / a[0]:10.1f}  {k}")


def report(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    for r in body:
        print("## " + r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
        wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tb = rd * scale.get(units[hdr.index("dram__bytes_read.sum")], 1) + wr * scale.get(units[hdr.index("dram__bytes_write.sum")], 1)
        print(f"  {'traffic = dram read + write (bytes per launch)':70s} {tb:16.0f}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1])
