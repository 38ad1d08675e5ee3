"""Key metrics of every launch in an `ncu --set full` report:  ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [out.txt]"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_bytes.sum", "L2 traffic"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(out=None):
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        lines.append(name[:160])
        for key, label in WANT:
            if key in idx:
                lines.append(f"    {label:26s} {r[idx[key]]:>16s} {units[idx[key]]}")
        lines.append("")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
