"""Prints value / e2e / per-kernel us of several bench.py JSON lines side by side.
usage: python tools/bench_compare.py gpurun_out/a.json gpurun_out/b.json ..."""
import json
import sys

rows = []
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    rows.append((f.split("/")[-1], d))
names = []
for _, d in rows:
    for k in d.get("kernels", {}):
        if k not in names:
            names.append(k)
print(f"{'':34s}" + "".join(f"{n[-22:]:>24s}" for n, _ in rows))
for key, fn in (("value Mpix/s", lambda d: d["value"]), ("ms/step", lambda d: d["ms_per_step"]),
                ("e2e Mpix/s", lambda d: d["e2e"]["value"]), ("sm MHz", lambda d: d["clocks"]["sm_mhz"])):
    print(f"{key:34s}" + "".join(f"{fn(d):24.3f}" for _, d in rows))
for k in names:
    print(f"{k + ' us':34s}" + "".join(f"{d['kernels'].get(k, {}).get('ms_per_launch', float('nan')) * 1e3:24.1f}" for _, d in rows))
