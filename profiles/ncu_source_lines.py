#!/usr/bin/env python3
"""Aggregate an ``ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K`` dump per CUDA
source line: share of executed instructions, share of stall samples and the main stall reasons.

    ncu -i gpurun_out/prof_r01_final.ncu-rep --page source --csv --print-source cuda,sass \\
        --kernel-name regex:detect_ewma > /tmp/detect.csv
    python profiles/ncu_source_lines.py /tmp/detect.csv [function-substring] [top-n]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    cur_file, cur_fn, hdr, ix, agg = None, "", None, {}, {}
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0] == "Line No":
            hdr, ix = r, {h: i for i, h in enumerate(r)}
        elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
            if want and want not in cur_fn[:80]:
                continue
            a = agg.setdefault((cur_file, int(r[0])), [0, 0, 0, 0, 0, 0, r[1].strip()[:78]])
            a[0] += int(r[ix["Instructions Executed"]] or 0)
            a[1] += int(r[ix["# Samples"]] or 0)
            for k, name in enumerate(("stall_long_sb", "stall_wait", "stall_short_sb", "stall_barrier")):
                a[2 + k] += int(r[ix[name]] or 0)
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print("instructions %d, stall samples %d; columns: inst%% samp%% (long_sb wait short_sb barrier, %% of all samples)" % (ti, ts))
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-16s %5d  %5.1f %5.1f  (%4.1f %4.1f %4.1f %4.1f)  %s" % (
            f[:16], ln, 100 * a[0] / ti, 100 * a[1] / ts, 100 * a[2] / ts, 100 * a[3] / ts, 100 * a[4] / ts, 100 * a[5] / ts, a[6]))


if __name__ == "__main__":
    main()
