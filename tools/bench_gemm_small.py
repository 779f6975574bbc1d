"""Back-to-back launch timing (host latency hidden) of tap-GEMM shapes with forced BLOCK_N variants. Writes gpurun_out/bench_gemm_small.json"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fatezero_b200 import ops
dev = "cuda"
REP = int(os.environ.get("REP", "20"))

def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(REP): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / REP

res = []
only = os.environ.get("ONLY")
for (M, N, K) in [(65536, 320, 320), (32768, 320, 320), (65536, 960, 320), (65536, 320, 1280), (16384, 640, 640), (16384, 1920, 640), (4096, 1280, 1280), (1024, 1280, 1280), (65536, 160, 320)]:
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for bn in ([0, 64, 128, 160, 256] if not only else [0]):
        if bn > N and bn != 0: continue
        ms = timeit(lambda: ops.gemm(a, w, out=out, force_bn=bn))
        res.append(dict(op="gemm", M=M, N=N, K=K, bn=bn, us=ms * 1e3, tflops=2 * M * N * K / ms / 1e9)); print(res[-1], flush=True)
    if only: break
if not only:
    for (M, N, K) in [(65536, 320, 320), (65536, 320, 1280), (16384, 640, 640), (4096, 1280, 1280)]:
        a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); out = torch.empty(M, N, device=dev, dtype=torch.float16)
        skip = torch.randn(M, N, device=dev).half(); bias = torch.randn(N, device=dev)
        ms = timeit(lambda: ops.gemm(a, w, bias=bias, residual=skip, out=out))
        print(dict(op="gemm+bias+residual", M=M, N=N, K=K, us=ms * 1e3, tflops=2 * M * N * K / ms / 1e9), flush=True)
    for (B, F, HW, Cin, Cout) in [(2, 8, 4096, 320, 160), (2, 8, 4096, 160, 320), (2, 8, 1024, 640, 160), (2, 8, 1024, 160, 640), (2, 8, 256, 1280, 160), (2, 8, 256, 160, 1280), (2, 8, 64, 1280, 160), (2, 8, 64, 160, 1280)]:
        x = torch.randn(B, F, HW, Cin, device=dev).half(); w3 = torch.randn(3, Cout, Cin, device=dev).half()
        ms = timeit(lambda: ops.tconv3(x, w3))
        res.append(dict(op="tconv3", HW=HW, Cin=Cin, Cout=Cout, us=ms * 1e3, tflops=2 * B * F * HW * Cin * Cout * 3 / ms / 1e9)); print(res[-1], flush=True)
    for (NB, H, Cin, Cout) in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280), (16, 8, 1280, 1280), (8, 64, 320, 320)]:
        x = torch.randn(NB, H, H, Cin, device=dev).half(); w9 = torch.randn(9, Cout, Cin, device=dev).half()
        for bn in [0, 128, 160, 256]:
            ms = timeit(lambda: ops.conv3x3(x, w9, force_bn=bn))
            res.append(dict(op="conv", NB=NB, H=H, Cin=Cin, Cout=Cout, bn=bn, us=ms * 1e3, tflops=2 * NB * H * H * Cin * Cout * 9 / ms / 1e9)); print(res[-1], flush=True)
    # host-side latency of one call (no sync): 
    import time
    a = torch.randn(128, 64, device=dev).half(); w = torch.randn(64, 64, device=dev).half(); out = torch.empty(128, 64, device=dev, dtype=torch.float16)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(200): ops.gemm(a, w, out=out)
    host_us = (time.perf_counter() - t) / 200 * 1e6
    torch.cuda.synchronize()
    res.append(dict(op="host_call_latency_us", us=host_us)); print(res[-1])
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_gemm_small.json"), "w"), indent=1)
