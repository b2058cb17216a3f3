#!/usr/bin/env python
"""Summarise ncu output (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/<tag>_launches.csv gpurun_out/<tag>_full.ncu-rep [profiles/ncu_traffic.json] > profiles/<name>.md
launch list -> per-kernel launch counts / total time / share; full capture -> key counters per kernel; the optional third
argument receives {kernel: DRAM bytes read + written per launch}, which bench.py reports as roofline.traffic."""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("cfm::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_val, i_met = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= i_val or r[i_met] != "gpu__time_duration.sum":
            continue
        k = short(r[i_name])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[i_val].replace(",", ""))
    unit = rows[1][hdr.index("Metric Unit")]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    tot = sum(v[1] for v in agg.values()) * scale
    print(f"## Launch list ({sum(v[0] for v in agg.values())} launches, {tot:.0f} us total, cold-cache serialised: compare shares)\n")
    print("| kernel | launches | total us | us/launch | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        t *= scale
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.3f} |")


METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe % (active)"),
    ("smsp__cycles_active.avg", "SMSP cycles active"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("\n## `ncu --set full` capture (one window of consecutive launches inside a step)\n")
    seen = set()
    traffic = {}
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        try:
            t = sum(float(r[idx[m]].replace(",", "")) * to_bytes[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            e = traffic.setdefault(k, {"dram_bytes_per_launch": 0.0, "launches": 0})
            e["dram_bytes_per_launch"] += t
            e["launches"] += 1
        except (KeyError, ValueError):
            pass
        if k in seen:
            continue
        seen.add(k)
        print(f"### `{k}`\n")
        for m, label in METRICS:
            if m in idx:
                print(f"* {label}: {r[idx[m]]} {units[idx[m]]}")
        print()
    for e in traffic.values():
        e["dram_bytes_per_launch"] /= max(1, e["launches"])
    return traffic


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        tr = full(sys.argv[2])
        if len(sys.argv) > 3:
            import json
            json.dump({"source": sys.argv[2], "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches of each kernel (ncu --set full, cold caches)",
                       "kernels": tr}, open(sys.argv[3], "w"), indent=1)
