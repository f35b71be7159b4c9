#!/usr/bin/env python
"""Warp-sample share per PHASE of k_brick from an `ncu --set full --import-source on` capture:
    python tools/ncu_phase_profile.py gpurun_out/r02b_k_brick.ncu-rep
Walks the SASS in address order with the CUDA source line of every instruction (--print-source sass + cuda correlation
is not in the CSV, so phases are cut at the BAR.SYNC / SYNCS instructions of the SASS stream) and sums samples and
executed instructions between them."""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr = None
    seg, segs = {"name": "start", "samples": 0, "inst": 0, "fp64": 0, "lds": 0, "ldg": 0}, []
    for row in rows:
        if not row: continue
        if row[0] == "Address": hdr = row; continue
        if hdr is None or len(row) < 8 or not row[0].startswith("0x"): continue
        sass = row[1].strip(); smp = int(row[hdr.index("# Samples")]); ins = int(row[hdr.index("Instructions Executed")])
        seg["samples"] += smp; seg["inst"] += ins
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        if op.startswith(("DFMA", "DADD", "DMUL", "DSETP", "MUFU.RCP64H", "DMNMX")): seg["fp64"] += ins
        if op.startswith("LDS"): seg["lds"] += ins
        if op.startswith(("LDG", "LDGSTS", "STG")): seg["ldg"] += ins
        if op.startswith(("BAR", "SYNCS", "WARPSYNC.ALL")) or "BAR.SYNC" in sass:
            segs.append(seg); seg = {"name": sass[:40], "samples": 0, "inst": 0, "fp64": 0, "lds": 0, "ldg": 0}
    segs.append(seg)
    tot = sum(s["samples"] for s in segs) or 1; ti = sum(s["inst"] for s in segs) or 1
    print("segment (ends with)                        samples%   inst%   fp64 inst  LDS  LDG/STG")
    for s in segs:
        print("%-42s %6.1f%% %6.1f%% %10d %8d %8d" % (s["name"], 100.0 * s["samples"] / tot, 100.0 * s["inst"] / ti, s["fp64"], s["lds"], s["ldg"]))
    print("total samples", tot, "warp instructions", ti)

if __name__ == "__main__":
    main()
