#!/usr/bin/env python
"""Turn ncu outputs under gpurun_out/ into the small text summaries kept under profiles/.

  launches:  tools/summarize_ncu.py launches <launches.csv> <out.txt>
  report:    tools/summarize_ncu.py report <file.ncu-rep> <out.txt>   (needs `ncu` on PATH)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"total_us {tot:.1f}\n")
        f.write(f"{'sum_us':>12} {'share':>7} {'n':>5} {'avg_us':>10}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:12.1f} {100 * t / tot:6.1f}% {n:5d} {t / n:10.1f}  {k}\n")


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({path}); selected raw metrics per captured launch\n")
        for data in rows[2:]:
            d, u = dict(zip(hdr, data)), dict(zip(hdr, units))
            f.write(f"\n## {d['Kernel Name'][:110]}  grid={d.get('launch__grid_size')} block={d.get('launch__block_size')}\n")
            for k in KEYS:
                if k in d and d[k] not in ("", "n/a"):
                    f.write(f"{k:75s} {d[k]:>16s} {u[k]}\n")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
