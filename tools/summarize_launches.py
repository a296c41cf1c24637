"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one ELBO step, per kernel."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
idx = [i for i, (n, _) in enumerate(names) if "fill_normal" in n]
starts = [i for k, i in enumerate(idx) if k == 0 or idx[k - 1] != i - 1]
start, end = starts[-2], starts[-1]
agg, tot = collections.OrderedDict(), 0.0
for n, t in names[start:end]:
    short = n.split("(")[0].replace("void ", "").replace("mogpe::", "")[:64]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
unit = rows[0]["Metric Unit"]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
print(f"one step: {end - start} launches, {tot * scale / 1e3:.3f} ms (ncu: serialised, cold cache)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t * scale:10.1f} us {100 * t / tot:5.1f}%  x{c:3d}  {k}")
