#!/usr/bin/env python
"""Summaries of ncu captures for profiles/ (run in the build container, where the .ncu-rep files land).

  python scripts/ncu_summary.py full  gpurun_out/prof_gemm.ncu-rep  > profiles/rNN_ncu_full_gemm.txt
  python scripts/ncu_summary.py list  gpurun_out/launches.csv       > profiles/rNN_ncu_launches.txt

`full`: one line per profiled launch with the metrics /opt/skills/guides/B200_PROFILING.md names (duration, DRAM
bytes read/written = `traffic`, DRAM %, tensor-pipe %, issue %, registers, L2 hit rate).
`list`: per-kernel totals of a `--metrics gpu__time_duration.sum` launch list, with each kernel's share.
"""
import collections
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full --clock-control none; one line per profiled launch")
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("regnet::<unnamed>::", "").replace("void ", "")
        parts = [f"{name:34s} grid={r[col['Grid Size']]:12s} block={r[col['Block Size']]:12s}"]
        for m, short in WANT:
            if m in col:
                parts.append(f"{short}={r[col[m]]}{units[col[m]]}")
        print(" ".join(parts))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    n = 0
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("regnet::<unnamed>::", "").replace("void ", "")[:64]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        c = tot.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += v
        n += 1
    s = sum(v[1] for v in tot.values())
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised); {n} launches, {s:.1f} us")
    for name, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print(f"{name:66s} {c:4d} launches {v:10.1f} us {100 * v / s:5.1f}%")


def traffic(path):
    """JSON for bench.py's roofline.traffic: DRAM bytes (read + write) of every profiled launch, and their sum."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, unit):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return float(v.replace(",", "")) * mult

    launches_ = []
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("regnet::<unnamed>::", "").replace("void ", "")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        t = r[col["gpu__time_duration.sum"]]
        launches_.append({"kernel": name, "dram_read_bytes": rd, "dram_write_bytes": wr,
                          "time": t + units[col["gpu__time_duration.sum"]]})
    print(json.dumps({"source": path, "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch",
                      "launches": launches_,
                      "dram_bytes_per_step": sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in launches_)}, indent=1))


if __name__ == "__main__":
    {"full": full, "list": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2])
