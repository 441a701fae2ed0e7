"""Summarise an ncu report (or a launch-list CSV) into the small text tables kept under profiles/.
usage: python tools/ncu_summary.py raw <report.ncu-rep>     |    launches <launches.csv>"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_%", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("lsu_%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("fma_%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("xu_%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("inst_M", "smsp__inst_executed.sum"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"{'kernel':34s}" + "".join(f"{n:>11s}" for n, _ in WANT) + "  top stalls (per issue)")
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("ts::", "")[:33]
        vals = []
        for n, m in WANT:
            v = float(r[idx[m]].replace(",", "")) if m in idx and r[idx[m]] else float("nan")
            u = units[idx[m]] if m in idx else ""
            if n == "time_us":
                v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            if n.endswith("_MB"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            if n == "inst_M":
                v = v / 1e6
            vals.append(v)
        st = sorted(((float(r[idx[h]]), h) for h in stall if r[idx[h]]), reverse=True)[:3]
        sts = ", ".join(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st)
        print(f"{name:34s}" + "".join(f"{v:11.2f}" for v in vals) + "  " + sts)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1e-3)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{n:72s} n={len(v):3d} avg_us={sum(v) / len(v):9.1f} share={sum(v) / tot:.3f}")


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2])
