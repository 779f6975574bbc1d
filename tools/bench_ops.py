"""Micro-benchmarks of the hot kernels on SD-1.4 shapes (CUDA events, L2 flushed between iterations).
Writes gpurun_out/bench_ops.json.  Usage: python tools/bench_ops.py [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fatezero_b200 import _lib, ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def _sel(only, name, items):
    return items if (not only or name in only) else []


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = []
    only = set(sys.argv[1:])  # e.g. `bench_ops.py norm` runs only the HBM-bound kernels
    BF = 16
    # linear layers (tokens x C)
    for (M, N, K) in _sel(only, "gemm", [(BF * 4096, 320, 320), (BF * 4096, 960, 320), (BF * 4096, 2560, 320), (BF * 4096, 320, 1280),
                      (BF * 1024, 640, 640), (BF * 1024, 5120, 640), (BF * 1024, 640, 2560), (BF * 256, 1280, 1280),
                      (BF * 256, 10240, 1280), (BF * 256, 1280, 5120), (8192, 8192, 8192)]):
        a = torch.randn(M, K, device=dev).half()
        w = torch.randn(N, K, device=dev).half()
        for bn in ([0] if N < 8192 else [256, 160, 128]):
            ms = timeit(lambda: ops.gemm(a, w, force_bn=bn))
            res.append(dict(op="gemm", M=M, N=N, K=K, bn=bn, ms=ms, tflops=2 * M * N * K / ms / 1e9))
            print(res[-1], flush=True)
    # convs
    for (NB, H, Cin, Cout, stride) in _sel(only, "conv", [(BF, 64, 320, 320, 1), (BF, 64, 640, 320, 1), (BF, 64, 960, 320, 1), (BF, 32, 640, 640, 1),
                                        (BF, 32, 1280, 640, 1), (BF, 32, 1920, 640, 1), (BF, 16, 1280, 1280, 1), (BF, 16, 2560, 1280, 1),
                                        (BF, 8, 1280, 1280, 1), (BF, 8, 2560, 1280, 1), (BF, 64, 320, 320, 2), (BF, 32, 640, 640, 2)]):
        x = torch.randn(NB, H, H, Cin, device=dev).half()
        w9 = torch.randn(9, Cout, Cin, device=dev).half()
        ms = timeit(lambda: ops.conv3x3(x, w9, stride=stride))
        fl = 2 * NB * (H // stride) ** 2 * Cin * Cout * 9
        res.append(dict(op="conv3x3", NB=NB, H=H, Cin=Cin, Cout=Cout, stride=stride, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)
    # attention
    for (S, d, heads, nsl, mode) in _sel(only, "attn", [(4096, 40, 8, 1, "none"), (1024, 80, 8, 1, "none"), (1024, 80, 8, 1, "store"),
                                      (1024, 80, 8, 1, "replace"), (256, 160, 8, 1, "store"), (64, 160, 8, 1, "store")]):
        C_ = heads * d
        q = torch.randn(BF * S, C_, device=dev).half()
        k = torch.randn(BF * S, C_, device=dev).half()
        vt = torch.randn(BF, heads, d, S, device=dev).half()
        out = torch.empty(BF * S, C_, device=dev, dtype=torch.float16)
        si = [[(b * 8 + 3) for b in range(BF // 8) for f in range(8)]] if nsl == 1 else None
        cache = torch.empty(BF, heads, S, S, device=dev, dtype=torch.float16) if mode != "none" else None
        if mode == "replace":
            cache.copy_(torch.softmax(torch.randn(BF, heads, S, S, device=dev), -1))
        kw = dict(S_q=S, keys_per_slot=S, n_src=BF, d=d, heads=heads, F=8, BF=BF, scale=d ** -0.5, src_index=si)
        if mode == "store":
            fn = lambda: ops.attention(q, k, vt, out, **kw, row_mode=_lib.ATTN_STORE, store=cache, cache_ld=S)
        elif mode == "replace":
            fn = lambda: ops.attention(q, k, vt, out, **kw, row_mode=_lib.ATTN_REPLACE, base=cache, cache_ld=S)
        else:
            fn = lambda: ops.attention(q, k, vt, out, **kw)
        ms = timeit(fn)
        fl = 4 * BF * heads * S * S * d
        res.append(dict(op="attn", S=S, d=d, mode=mode, ms=ms, tflops=fl / ms / 1e9,
                        cache_gbs=(BF * heads * S * S * 2 / ms / 1e6) if mode != "none" else None))
        print(res[-1], flush=True)
    # groupnorm
    for (NB, HW, Cc) in _sel(only, "norm", [(BF, 4096, 320), (BF, 4096, 960), (BF, 1024, 1280), (BF, 256, 2560)]):
        x = torch.randn(NB, HW, Cc, device=dev).half()
        g = torch.ones(Cc, device=dev)
        b = torch.zeros(Cc, device=dev)
        ms = timeit(lambda: ops.groupnorm(x, g, b, 1e-5, 32, 8, True))
        res.append(dict(op="groupnorm", NB=NB, HW=HW, C=Cc, ms=ms, gbs=NB * HW * Cc * 2 * 3 / ms / 1e6))
        print(res[-1], flush=True)
    # temporal attention + layernorm
    for (B, HW, heads, d) in _sel(only, "norm", [(2, 4096, 8, 40), (2, 1024, 8, 80), (2, 256, 8, 160)]):
        C_ = heads * d
        qkv = torch.randn(B * 8 * HW, 3 * C_, device=dev).half()
        ms = timeit(lambda: ops.temporal_attn(qkv, B, 8, HW, heads, d, d ** -0.5))
        res.append(dict(op="temporal_attn", B=B, HW=HW, C=C_, ms=ms, gbs=B * 8 * HW * C_ * 2 * 4 / ms / 1e6))
        print(res[-1], flush=True)
        x = torch.randn(B * 8 * HW, C_, device=dev).half()
        g = torch.ones(C_, device=dev)
        b = torch.zeros(C_, device=dev)
        ms = timeit(lambda: ops.layernorm(x, g, b))
        res.append(dict(op="layernorm", M=B * 8 * HW, C=C_, ms=ms, gbs=B * 8 * HW * C_ * 2 * 2 / ms / 1e6))
        print(res[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
