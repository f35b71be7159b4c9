#!/usr/bin/env python
"""Warp-sample profile per CUDA source line from an `ncu --set full --import-source on` capture (read here, no GPU):

    python tools/ncu_line_profile.py gpurun_out/k_elem_affine.ncu-rep [top=30]

Prints the share of the stall samples per source line with the dominant stall reasons (L long scoreboard, W wait,
M math-pipe throttle, NS not selected, S selected, sh short scoreboard) -- the table behind DESIGN.md section 3.11."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, cur, agg = None, "", []
    for row in rows:
        if not row:
            continue
        if row[0] in ("File Name", "Function Name"):
            cur = row[1]
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr and len(row) > 7 and row[2] == "-":
            try:
                agg.append((cur.split("/")[-1], int(row[0]), row[1].strip(), int(row[4]), row))
            except ValueError:
                pass
    tot = sum(a[3] for a in agg) or 1
    idx = {k: hdr.index(k) for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_short_sb")}
    print("total samples", tot)
    for f, ln, src, n, row in sorted(agg, key=lambda x: -x[3])[:top]:
        print("%5.1f%% %6d  L%-5s W%-5s M%-5s NS%-5s S%-5s sh%-4s %s:%d | %s" % (
            100.0 * n / tot, n, row[idx["stall_long_sb"]], row[idx["stall_wait"]], row[idx["stall_math"]], row[idx["stall_not_selected"]],
            row[idx["stall_selected"]], row[idx["stall_short_sb"]], f[:24], ln, src[:90]))


if __name__ == "__main__":
    main()
