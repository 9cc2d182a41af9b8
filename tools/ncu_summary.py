#!/usr/bin/env python3
"""Text summary of the round's ncu evidence (what gets committed under profiles/; the .ncu-rep files stay in gpurun_out/).

  python tools/ncu_summary.py <launches.csv> <stages.ncu-rep> [title] > profiles/rN_stages_ncu_summary.txt

  launches.csv     ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>
  stages.ncu-rep   ncu --set full --clock-control none --import-source on -k regex:'k_extend_pre|k_traverse|k_extend_post|k_shade'
                   -s 9 -c 4 -o stages <cmd>      (launches 9..12 = the third full iteration of the wavefront)
Needs `ncu` on PATH (reads the report here, no GPU involved).
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors.sum.per_second",
    "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launch_list(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr_i]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = []
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"  {name:28s} launches={n:4d} total={t / 1e6:8.3f} ms share={t / tot:.3f} avg={t / n / 1e3:8.1f} us")
    return out


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kernels = rows[2:]
    ki = hdr.index("Kernel Name")
    cols = {}
    for r in kernels:
        cols[r[ki].split("(")[0]] = dict(zip(hdr, r))
    return cols, dict(zip(hdr, units))


def main():
    launches, rep = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else ""
    print(title)
    print("\n== launch list: ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised: compare SHARES) ==")
    print("\n".join(launch_list(launches)))
    cols, units = raw_page(rep)
    names = list(cols)
    print("\n== ncu --set full --clock-control none, one launch of each stage kernel (third full iteration) ==")
    print("  " + "metric".ljust(72) + " ".join(n[:17].ljust(17) for n in names))
    for m in METRICS:
        print("  " + f"{m} [{units.get(m, '')}]".ljust(72) + " ".join(str(cols[n].get(m, "-"))[:17].ljust(17) for n in names))
    stall = [h for h in cols[names[0]] if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             and "not_issued" not in h]
    for h in sorted(stall):
        short = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
        print("  " + f"stall:{short} [warps per issue-active cycle]".ljust(72) + " ".join(str(cols[n].get(h, "-"))[:17].ljust(17) for n in names))
    traffic = {n: (float(cols[n]["dram__bytes_read.sum"].replace(",", "")), float(cols[n]["dram__bytes_write.sum"].replace(",", "")),
                   units.get("dram__bytes_read.sum", "")) for n in names}
    print("\n== DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) ==")
    print("  " + json.dumps({n: {"read": t[0], "write": t[1], "unit": t[2]} for n, t in traffic.items()}))


if __name__ == "__main__":
    main()
