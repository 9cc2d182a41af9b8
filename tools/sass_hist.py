#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libptcore.so: the checkable form of the Blackwell-specific claims in DESIGN.md
(packed FP32 `FFMA2`, 16-byte read-only loads `LDG.E.128.CONSTANT`, 3-input `FMNMX3`, byte-permute node decode `PRMT`,
reciprocal / sqrt seeds `MUFU.*`, and what the per-lane traversal stack costs in local-memory `LDL` / `STL`).

  cuobjdump -sass raytracer-rust_b200/libptcore.so | python tools/sass_hist.py > profiles/r2_sass_hist.txt
"""
import collections
import re
import sys

FOCUS = ["FFMA2", "FFMA", "FMUL", "FADD", "FMNMX3", "FMNMX", "MUFU.RCP", "MUFU.RSQ", "MUFU.SQRT", "PRMT", "LOP3", "SHF", "SEL",
         "LDG.E.128.CONSTANT", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS", "LDL", "STL",
         "ATOMS", "ATOMG", "RED", "VOTE", "SHFL", "MATCH", "BAR", "BRA", "CALL", "ACQBULK", "UBLKCP"]


def main():
    kern, hist, full = None, collections.OrderedDict(), {}
    pat = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)")
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            full[kern] = collections.Counter()
            continue
        m = pat.match(line)
        if m and kern:
            op = m.group(1)
            full[kern][op] += 1
            hist[kern][op.split(".")[0]] += 1
    print("# SASS opcode histogram per kernel (static instruction counts, sm_100a), from `cuobjdump -sass libptcore.so`")
    for k, h in hist.items():
        name = re.sub(r"^_ZN3ptw\d+", "", k)
        total = sum(h.values())
        print(f"\n== {k}\n   {total} instructions")
        row = []
        for f in FOCUS:
            n = sum(c for op, c in full[k].items() if op == f or op.startswith(f + ".")) if "." in f else h.get(f, 0)
            if n:
                row.append(f"{f}={n}")
        print("   focus: " + "  ".join(row))
        print("   top:   " + "  ".join(f"{op}={c}" for op, c in h.most_common(14)))
        _ = name


if __name__ == "__main__":
    main()
