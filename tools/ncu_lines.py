#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line.

  ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > prof_cs.csv ; python tools/ncu_lines.py prof_cs.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur, hdr, agg = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not (hdr and cur and len(r) == len(hdr)) or r[0] == "":
        continue
    i_smp, i_ie, i_te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    try:
        ie, te, smp = float(r[i_ie]), float(r[i_te]), float(r[i_smp])
    except ValueError:
        continue
    stalls = {h: float(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v not in ("", "-")}
    agg.append((ie, te, smp, cur.split("/")[-1], r[0], r[1].strip()[:90], stalls))
tot, tots = sum(a[0] for a in agg), sum(a[2] for a in agg)
print(f"total warp instructions {tot:.0f}, thread instructions {sum(a[1] for a in agg):.0f} "
      f"(avg active threads {sum(a[1] for a in agg) / tot:.2f}), samples {tots:.0f}")
by_file = collections.defaultdict(lambda: [0, 0, 0])
for a in agg:
    by_file[a[3]][0] += a[0]
    by_file[a[3]][1] += a[1]
    by_file[a[3]][2] += a[2]
for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:18s} inst share {v[0] / tot:.3f}  avg threads {v[1] / max(v[0], 1):5.1f}  sample share {v[2] / tots:.3f}")
st = collections.Counter()
for a in agg:
    st.update(a[6])
print("  stall samples:", ", ".join(f"{k[6:]} {v / sum(st.values()):.2f}" for k, v in st.most_common(8)))
for a in sorted(agg, key=lambda a: -a[2])[:top]:
    s = ",".join(f"{k[6:]}:{int(v)}" for k, v in sorted(a[6].items(), key=lambda kv: -kv[1])[:3] if v > 0)
    print(f"{100 * a[0] / tot:6.2f}%i {100 * a[2] / tots:6.2f}%s thr={a[1] / max(a[0], 1):4.1f} {a[3]}:{a[4]:>4s}  {a[5]}   [{s}]")
