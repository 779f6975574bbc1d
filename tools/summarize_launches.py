"""Summarise an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]) per kernel name.
  python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import io
import re
import sys


def main():
    path = sys.argv[1]
    txt = open(path, errors="replace").read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void (fz::)?", "", name)
        if "at::" in name or "vectorized_elementwise" in name or "elementwise_kernel" in name or "CatArray" in name:
            name = "torch elementwise (latent copies / casts)"
        d = agg.setdefault(name, dict(ids=set(), t=0.0, rd=0.0, wr=0.0))
        d["ids"].add(r["ID"])
        v = float(r["Metric Value"].replace(",", ""))
        m, u = r["Metric Name"], r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            d["t"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[u]
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            d["rd" if "read" in m else "wr"] += v * scale
    tot = sum(d["t"] for d in agg.values())
    n = sum(len(d["ids"]) for d in agg.values())
    lines = ["| kernel | launches | time (ms) | share | DRAM read (GB) | DRAM write (GB) | DRAM GB/s while running |", "|---|---|---|---|---|---|---|"]
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = (d["rd"] + d["wr"]) / 1e9 / max(d["t"] / 1e3, 1e-12)
        lines.append(f"| `{name}` | {len(d['ids'])} | {d['t']:.2f} | {100 * d['t'] / tot:.1f} % | {d['rd'] / 1e9:.2f} | {d['wr'] / 1e9:.2f} | {gbs:.0f} |")
    lines.append("")
    lines.append(f"total: {n} launches, {tot:.2f} ms")
    out = "\n".join(lines)
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(out + "\n")


if __name__ == "__main__":
    main()
