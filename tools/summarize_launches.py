"""Summarise an `ncu --csv --metrics ...` launch list: per kernel name -> count, avg time, avg cycles, tensor %.

    python tools/summarize_launches.py gpurun_out/launches.csv
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, ni, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("vrag::", "").replace("<unnamed>::", "")
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        per.setdefault((r[idi], name), {})[r[ni]] = v
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, {"n": 0, "t": 0.0, "c": 0.0, "tp": 0.0})
        a["n"] += 1
        a["t"] += m.get("gpu__time_duration.sum", 0.0)
        a["c"] += m.get("sm__cycles_elapsed.max", 0.0)
        a["tp"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    tot = sum(a["t"] for a in agg.values()) or 1.0
    print(f"{'kernel':58s} {'n':>5s} {'avg us':>9s} {'avg kcyc':>9s} {'tensor%':>8s} {'share':>6s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        n = a["n"]
        print(f"{k[:58]:58s} {n:5d} {a['t'] / n / 1e3:9.1f} {a['c'] / n / 1e3:9.1f} {a['tp'] / n:8.1f} {a['t'] / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
