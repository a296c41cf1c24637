"""profiles/traffic.json from an `ncu --set full` capture: dram read + write bytes PER LAUNCH of the named kernels.

    ncu -i gpurun_out/step.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv profiles/<summary>.csv KEY=substring[:algorithmic_bytes] ...

For every KEY the launches whose kernel name contains `substring` are averaged (the largest one is also recorded).  bench.py
reads profiles/traffic.json for `roofline.traffic`; it never prints a constant typed into the source.
"""
import csv
import json
import os
import sys


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    raw, source = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    rd, wr = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for spec in sys.argv[3:]:
        key, rest = spec.split("=", 1)
        sub, _, alg = rest.partition(":")
        sel = [r for r in body if sub in r[col["Kernel Name"]]]
        if not sel:
            print(f"{key}: no launch matches '{sub}'")
            continue
        per = [to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]) for r in sel]
        out[key] = {"dram_bytes_per_launch": sum(per) / len(per), "largest_launch_bytes": max(per), "launches": len(per),
                    "what": sub, "source": source, "algorithmic_bytes": float(alg) if alg else None}
        print(key, out[key])
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
