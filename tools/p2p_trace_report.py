#!/usr/bin/env python
"""Phase breakdown of the partitioned step from the FTB200_P2P_TRACE dumps (one file per rank).

    python tools/p2p_trace_report.py gpurun_out/r02q_trace [first_step]

Slots (ns, %globaltimer of each rank's GPU): 0 step start | 1 boundary elements done | 2 packed + flagged |
3 k_adv_p2p starts (interior joined) | 4 dt published | 5 all ranks' dt + neighbours' flags seen | 6 node kernel done |
7 interior elements done.  Prints medians in microseconds over the steps >= first_step."""
import glob
import sys

import numpy as np


def main():
    prefix = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    for path in sorted(glob.glob(prefix + "_rank*.txt")):
        a = np.loadtxt(path, dtype=np.int64)
        a = a[a[:, 0] >= first]
        t = a[:, 1:].astype(np.float64) * 1e-3
        ok = np.all(t > 0, axis=1)
        t = t[ok]
        step = np.diff(t[:, 0])
        seg = {
            "step (start to next start)": step,
            "boundary elements (0->1)": t[:, 1] - t[:, 0],
            "pack + flags (1->2)": t[:, 2] - t[:, 1],
            "interior elements done (0->7)": t[:, 7] - t[:, 0],
            "join -> k_adv_p2p starts (7->3)": t[:, 3] - t[:, 7],
            "publish dt (3->4)": t[:, 4] - t[:, 3],
            "wait for peers (4->5)": t[:, 5] - t[:, 4],
            "adv tail + node kernel (5->6)": t[:, 6] - t[:, 5],
            "node done -> next step start (6->0')": t[1:, 0] - t[:-1, 6],
        }
        print(path, "(%d steps)" % len(t))
        for k, v in seg.items():
            print("  %-40s median %8.2f  p10 %8.2f  p90 %8.2f us" % (k, np.median(v), np.percentile(v, 10), np.percentile(v, 90)))


if __name__ == "__main__":
    main()
