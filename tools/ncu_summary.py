#!/usr/bin/env python
"""Turn an .ncu-rep capture into the flat `metric,unit,value` CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/k_elem_affine.ncu-rep profiles/r01_k_elem_affine_ncu_full.csv

Reads the report with `ncu -i REP --page raw --csv` (the longest captured launch) and writes one line per metric."""
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr = next(i for i, row in enumerate(rows) if row and row[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    data = [r for r in rows[hdr + 2:] if len(r) == len(names)]
    # several captured launches (e.g. the variants of k_node): keep the longest one -- the main kernel of the step
    ti = names.index("gpu__time_duration.sum") if "gpu__time_duration.sum" in names else None
    vals = max(data, key=lambda r: float(r[ti].replace(",", ""))) if ti is not None else data[0]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for n, u, v in zip(names, units, vals):
            if n in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
                continue
            w.writerow([n, u, v])
    print("wrote %s (%d metrics)" % (out, len(names)))


if __name__ == "__main__":
    main()
