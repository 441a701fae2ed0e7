"""Per-region breakdown (stall samples, executed instructions) of one kernel from an ncu source-page CSV:
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
  python tools/ncu_regions.py src.csv off1:name1 off2:name2 ...   (SASS byte offsets where each region starts)"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    idx = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    regions = sorted((int(a.split(":")[0], 16), a.split(":")[1]) for a in sys.argv[2:]) or [(0, "all")]
    base = int(data[0][0], 16)

    def region(off):
        name = regions[0][1]
        for o, n in regions:
            if off >= o:
                name = n
        return name

    tot = sum(int(r[idx["# Samples"]]) for r in data)
    reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[idx[n]]) for r in data) for n in reasons}
    print("total samples", tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {k:26s} {v:8d} {v / tot:.3f}")
    rs, ri, rt = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        reg = region(int(r[0], 16) - base)
        rs[reg] += int(r[idx["# Samples"]])
        ri[reg] += int(r[idx["Instructions Executed"]])
        rt[reg] += int(r[idx["Predicated-On Thread Instructions Executed"]])
    for _, name in regions:
        lanes = rt[name] / max(ri[name], 1)
        print(f"{name:12s} samples {rs[name]:8d} {rs[name] / tot:.3f}   warp-inst {ri[name] / 1e6:8.1f} M   active lanes/inst {lanes:5.1f}")


if __name__ == "__main__":
    main()
