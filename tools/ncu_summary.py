"""Summarise an ncu report (.ncu-rep) or a `--metrics gpu__time_duration.sum --csv` launch list into the
short text tables kept under profiles/.

    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep  > profiles/rNN_<kernel>_full.txt
    python tools/ncu_summary.py list  gpurun_out/launches.csv  > profiles/rNN_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__cycles_active.avg", "sm__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1], rows[start + 2:]


def full(rep):
    hdr, units, rows = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        if len(r) < len(hdr):
            continue
        print(f"== {r[col['Kernel Name']]}  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}")
        for k in KEYS:
            hits = [h for h in hdr if h == k or h.endswith("." + k) or h.endswith(k)]
            for h in hits[:1]:
                print(f"   {k:78s} {r[col[h]]:>16s} {units[col[h]]}")
        print()


def launch_list(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]].split("(")[0]
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        us = v / 1e3 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
        total += us; n += 1
    print(f"{n} launches, {total:.1f} us total (ncu per-launch times are cold-cache and serialised: compare shares)")
    print(f"{'kernel':90s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:90]:90s} {cnt:8d} {us:10.1f} {us / cnt:9.2f} {100 * us / total:6.1f}%")


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    full(path) if mode == "full" else launch_list(path)
