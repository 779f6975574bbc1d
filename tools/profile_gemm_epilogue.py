"""Cycle accounting of the tap-GEMM warps (library built with -DFZ_GEMM_PROFILE: FZ_LIB_VARIANT=gprof) for a few memory-bound shapes."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fatezero_b200 import _lib, ops
dev = "cuda"
lib = _lib.load()
NAMES = ["epi total", "wait tfull", "tmem ld", "math", "wait slot", "stage+store", "chunks", "tiles", "mma total", "mma wait tempty",
         "mma wait full", "-", "prod total", "prod wait empty", "-", "-"]

def report(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    lib.fz_debug_gemm_counters(buf)
    v = list(buf)
    print(f"{name}: {s.elapsed_time(e) * 1e3:.1f} us  " + ", ".join(f"{n}={x}" for n, x in zip(NAMES, v) if n != "-"), flush=True)

a = torch.randn(65536, 320, device=dev).half()
w = torch.randn(320, 320, device=dev).half() * 0.05
bias = torch.zeros(320, device=dev)
res = torch.randn(65536, 320, device=dev).half()
report("linear 65536x320x320", lambda: ops.gemm(a, w))
report("linear 65536x320x320 +bias +skip", lambda: ops.gemm(a, w, bias=bias, residual=res))
wq = torch.randn(960, 320, device=dev).half() * 0.05
report("linear 65536x960x320", lambda: ops.gemm(a, wq))
wg = torch.randn(2560, 320, device=dev) * 0.05
bg = torch.zeros(2560, device=dev)
wp, bp, bn = ops.pack_geglu(wg, bg)
wp = wp.half()
report("GEGLU 65536x2560x320", lambda: ops.gemm(a, wp, bias=bp, geglu=True, force_bn=bn))
a2 = torch.randn(16384, 640, device=dev).half(); w2 = torch.randn(640, 640, device=dev).half() * 0.05
report("linear 16384x640x640", lambda: ops.gemm(a2, w2))
x = torch.randn(16, 64, 64, 320, device=dev).half(); w9 = torch.randn(9, 320, 320, device=dev).half() * 0.02
report("conv3x3 320->320 @64 BF=16", lambda: ops.conv3x3(x, w9))
