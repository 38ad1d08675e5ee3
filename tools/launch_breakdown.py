"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
for ONE learner step (from one prep_rows_kernel launch to the next)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "prep_rows" in n]
    i0, i1 = starts[0], starts[1]
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[i0:i1]:
        v = float(r["Metric Value"].replace(",", "")) / 1e3
        key = re.sub(r"^void ", "", r["Kernel Name"])
        key = re.sub(r"\(.*", "", key)[:150]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    lines = [f"one step: {i1 - i0} launches, sum of kernel durations {tot:.1f} us (ncu, cold-cache, serialised)"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{a[1]:9.1f} us {a[0]:4d} {100 * a[1] / tot:5.1f}%  {k}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
