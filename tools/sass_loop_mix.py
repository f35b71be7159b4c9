#!/usr/bin/env python
"""Instruction mix of a kernel and of its largest loop, from the SASS of the built library (no GPU needed).

    python tools/sass_loop_mix.py femtech_b200/libftb200.so k_elem_affineILi1ELb0
    python tools/sass_loop_mix.py femtech_b200/libftb200.so 6k_elemILi1ELb1ELb1ELb0

The second argument is a substring of the mangled kernel name.  The "loop" is the span of the longest backward branch
(the Gauss-point loop of the element kernels); DFMA + DADD + DMUL there is the fp64-pipe cost per Gauss point quoted in
DESIGN.md (169 for k_elem_affine, 204-205 for k_elem)."""
import re
import subprocess
import sys

OPS = ["DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDS", "STS", "LDL", "STL", "LDG", "STG", "LDGSTS", "IMAD", "LOP3", "MOV", "FSEL",
       "ISETP", "BRA", "CALL"]


def mix(sel):
    c = {}
    for _, t in sel:
        t = re.sub(r"^@!?U?P\d\s+", "", t)
        op = t.split()[0].split(".")[0]
        c[op] = c.get(op, 0) + 1
    return c


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = sorted(set(m for m in re.findall(r"Function : (\S+)", names) if pat in m))
    for fn in funcs:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
        ins = []
        for l in out.splitlines():
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        best = None
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                    best = (tgt, a)
        allc = mix(ins)
        print(fn)
        print("  whole kernel: %d instructions" % len(ins), {k: allc[k] for k in OPS if k in allc})
        if best:
            loop = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
            lc = mix(loop)
            print("  largest loop %#x-%#x: %d instructions" % (best[0], best[1], len(loop)), {k: lc[k] for k in OPS if k in lc})
            print("  fp64-pipe instructions per trip: %d" % sum(lc.get(k, 0) for k in ("DFMA", "DADD", "DMUL")))


if __name__ == "__main__":
    main()
