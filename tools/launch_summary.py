"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: per-kernel totals of
one fullProve step (the span between two witness-kernel launches)."""
import collections
import csv
import sys


def main(path, step=3):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    idx = [i for i, n in enumerate(names) if "witness" in n.lower()]
    seg = rows[idx[step]:idx[step + 1]]
    d = collections.OrderedDict()
    tot = 0
    for x in seg:
        k = x["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("nzcb::", "")[:50]
        v = float(x["Metric Value"]) / 1e6
        a = d.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"step launches {len(seg)}  total {tot:.3f} ms (sum of per-launch gpu__time_duration, serialised, cold cache)")
    for k, (c, v) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:52s} {c:4d} {v:9.3f} ms {100 * v / tot:5.1f}%  avg {v / c:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
