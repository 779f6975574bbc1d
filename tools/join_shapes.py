"""Join an ncu launch list of tools/profile_step.py with the shape sequence it recorded (gpurun_out/profile_step_shapes.json): the i-th
tapgemm / attention kernel of the list is the i-th recorded launch.  Prints per-shape GPU time, TFLOP/s and algorithmic GB/s.
  python tools/join_shapes.py gpurun_out/launches.csv gpurun_out/profile_step_shapes.json [out.json]"""
import csv
import io
import json
import re
import sys


def main():
    txt = open(sys.argv[1], errors="replace").read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.find('"ID"'):])))
    seq = json.load(open(sys.argv[2]))
    dur = []
    seen = set()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum" or r["ID"] in seen:
            continue
        name = r["Kernel Name"]
        if "temporal" not in name and ("tapgemm" in name or "attn_kernel" in name or "attn_plain" in name or "attn_cross" in name):
            seen.add(r["ID"])
            v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}[r["Metric Unit"]]
            dur.append((re.sub(r"\(.*", "", name).replace("void ", "").replace("fz::", ""), v))
    assert len(dur) == len(seq), (len(dur), len(seq))
    agg = {}
    for (kname, us), (ep, key) in zip(dur, seq):
        if ep == "gemm":
            (M, K), (N, _), geglu, res = key
            flops, byts = 2.0 * M * N * K, 2.0 * (M * K + N * K + M * (N // 2 if geglu else N) + (M * N if res else 0))
        elif ep == "conv3x3":
            (NB, H, W, Ci), (_, Co, _), stride = key
            rows_ = NB * (H // stride) * (W // stride)
            flops, byts = 2.0 * rows_ * Co * Ci * 9, 2.0 * (NB * H * W * Ci + 9 * Co * Ci + rows_ * Co)
        elif ep == "tconv3":
            (B, F, HW, Ci), (_, Co, _) = key
            flops, byts = 2.0 * B * F * HW * Co * Ci * 3, 2.0 * (B * F * HW * (Ci + Co) + 3 * Co * Ci)
        else:
            S, T, d, BF, mode, start = key
            flops, byts = 4.0 * BF * 8 * S * T * d, 0.0
        k = f"{ep}{key} [{kname}]"
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += flops; a[3] += byts
    tot = sum(a[1] for a in agg.values())
    out = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(dict(launch=k, n=a[0], us_total=round(a[1], 1), share=round(a[1] / tot, 4), us_each=round(a[1] / a[0], 1),
                        tflops=round(a[2] / a[1] / 1e6, 1), alg_gbs=round(a[3] / a[1] / 1e3, 0)))
    for o in out[:60]:
        print(f"{o['us_total']:9.1f} us {100 * o['share']:5.1f}%  n={o['n']:3d}  {o['us_each']:8.1f} us  {o['tflops']:7.1f} TF  {o['alg_gbs']:6.0f} GB/s  {o['launch']}")
    print("total", round(tot / 1e3, 2), "ms over", len(dur), "launches")
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
