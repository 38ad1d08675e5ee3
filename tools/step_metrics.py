"""Per-kernel table of ONE learner step from an ncu CSV that holds several metrics per launch, e.g.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file L.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline

usage: step_metrics.py L.csv [out.txt] [traffic.json]
The step is the launch range between two consecutive prep_rows_kernel launches.  traffic.json receives the DRAM bytes of
the K1+K1b unroll group (every tcgen05 GEMM/conv kernel and both recurrence kernels) -- bench.py's roofline.traffic."""
import collections
import csv
import json
import re
import sys

UNROLL = ("umma2_kernel", "umma3_kernel", "winconv_kernel", "winwgrad_kernel", "rec_fwd_kernel", "rec_bwd_kernel", "rec2_fwd_kernel", "rec2_bwd_kernel")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, out=None, traffic=None):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v, unit = num(r["Metric Value"]), r.get("Metric Unit", "")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e6, "us": 1e3, "s": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = v * scale
    ls = list(launches.values())
    starts = [i for i, d in enumerate(ls) if "prep_rows" in d["name"]]
    step = ls[starts[0]:starts[1]]
    agg = collections.OrderedDict()
    for d in step:
        key = re.sub(r"^void ", "", d["name"])
        key = re.sub(r"\(.*", "", key)[:120]
        a = agg.setdefault(key, collections.Counter())
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["l2"] += d.get("lts__t_bytes.sum", 0.0)
        a["tw"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    tot = collections.Counter()
    for a in agg.values():
        tot.update(a)
    lines = [f"one step ({len(step)} launches): {tot['us']:.0f} us (ncu: cold-cache, serialised), DRAM read {tot['rd'] / 1e9:.2f} GB, "
             f"write {tot['wr'] / 1e9:.2f} GB, L2 traffic {tot['l2'] / 1e9:.2f} GB",
             "      us   n DRAM rd MB    wr MB   L2 GB DRAM GB/s L2 TB/s tensor%  kernel"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
        t = a["us"] * 1e-6
        lines.append(f"{a['us']:8.1f} {a['n']:3d} {a['rd'] / 1e6:10.1f} {a['wr'] / 1e6:8.1f} {a['l2'] / 1e9:7.2f} "
                     f"{(a['rd'] + a['wr']) / t / 1e9 if t else 0:9.0f} {a['l2'] / t / 1e12 if t else 0:7.2f} "
                     f"{a['tw'] / (a['us'] * 1e3) if a['us'] else 0:7.1f}  {k}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")
    if traffic:
        sel = [d for d in step if any(u in d["name"] for u in UNROLL)]
        b = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in sel)
        json.dump({"what": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the K1+K1b unroll-group launches (" + ", ".join(UNROLL) +
                           ") of one learner step, C=4, strict precision, ncu --clock-control none",
                   "bytes": b, "launches": len(sel), "sum_us": sum(d.get("gpu__time_duration.sum", 0.0) for d in sel) / 1e3},
                  open(traffic, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
