"""Summarise an ncu report (`ncu --set full`) into a markdown table of the counters the design argues with.
  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_summary.py /tmp/raw.csv [labels.txt] > profiles/xxx.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU pipe %"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    labels = [l.strip() for l in open(sys.argv[2])] if len(sys.argv) > 2 else None
    idx = {n: i for i, n in enumerate(names)}
    print("| launch | kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for j, r in enumerate(rows[hdr + 2:]):
        if not r or len(r) < len(names):
            continue
        kname = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("fz::", "")
        vals = []
        for key, _ in COLS:
            if key in idx:
                v, u = r[idx[key]], units[idx[key]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.1f}" if abs(f) < 1e5 else f"{f:.3g}"
                except ValueError:
                    pass
                vals.append(f"{v} {u}".strip())
            else:
                vals.append("-")
        lab = labels[j] if labels and j < len(labels) else str(j)
        print(f"| {lab} | `{kname}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
