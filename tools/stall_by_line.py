#!/usr/bin/env python
"""Warp-stall samples of one captured launch aggregated by CUDA source line (needs `ncu`, no GPU):

    python tools/stall_by_line.py <file.ncu-rep> <kernel regex> <launch skip> [top N] > profiles/....txt

Uses `ncu --page source --print-source cuda,sass --csv` (the capture must have been taken with --import-source on and
the library built with -lineinfo)."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, regex, skip = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + regex, "--launch-skip", skip, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file, hdr, func = None, None, None
    per_line = collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    text = {}
    total = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue  # SASS rows (no line number) are already folded into their source line
        d = dict(zip(hdr, r))
        try:
            n = int(d["# Samples"])
        except ValueError:
            continue
        if n == 0:
            continue
        key = (cur_file, int(r[0]))
        per_line[key] += n
        total += n
        text[key] = r[1].strip()[:86]
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                stalls[key][k[6:]] += int(v)
    print(f"# {rep}: {func}")
    print(f"# warp-stall samples by source line (all warps of the CTA: 4 control + 8 epilogue), total {total}")
    for key, n in per_line.most_common(top):
        why = ", ".join(f"{k} {v}" for k, v in stalls[key].most_common(2))
        print(f"{100 * n / max(total, 1):5.1f}%  {key[0]}:{key[1]:<4d} {text[key]:86s} [{why}]")


if __name__ == "__main__":
    main()
