"""One launch each of the HBM-bound kernels for `ncu --set full` (inputs produced right before, i.e. L2-warm like in the step):
GroupNorm+SiLU 16x4096x320 (joint over 8 frames), GroupNorm 16x4096x960, LayerNorm 65536x320, temporal attention r=64 / r=32 / r=16,
plus small tap-GEMMs (to_out 65536x320x320 with bias+residual, r=8 conv 1280->1280)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fatezero_b200 import ops
dev = "cuda"
x320 = torch.randn(16, 4096, 320, device=dev).half(); x960 = torch.randn(16, 4096, 960, device=dev).half()
g320, b320 = torch.ones(320, device=dev), torch.zeros(320, device=dev)
g960, b960 = torch.ones(960, device=dev), torch.zeros(960, device=dev)
xl = torch.randn(65536, 320, device=dev).half()
qkv64 = torch.randn(2 * 8 * 4096, 960, device=dev).half(); qkv32 = torch.randn(2 * 8 * 1024, 1920, device=dev).half()
qkv16 = torch.randn(2 * 8 * 256, 3840, device=dev).half()
a = torch.randn(65536, 320, device=dev).half(); w = torch.randn(320, 320, device=dev).half() * 0.05; bias = torch.zeros(320, device=dev)
res = torch.randn(65536, 320, device=dev).half()
x8 = torch.randn(16, 8, 8, 1280, device=dev).half(); w8 = torch.randn(9, 1280, 1280, device=dev).half() * 0.01
fns = [lambda: ops.groupnorm(x320, g320, b320, 1e-5, 32, 8, True), lambda: ops.groupnorm(x960, g960, b960, 1e-5, 32, 8, True),
       lambda: ops.layernorm(xl, g320, b320), lambda: ops.temporal_attn(qkv64, 2, 8, 4096, 8, 40, 40 ** -0.5),
       lambda: ops.temporal_attn(qkv32, 2, 8, 1024, 8, 80, 80 ** -0.5), lambda: ops.temporal_attn(qkv16, 2, 8, 256, 8, 160, 160 ** -0.5),
       lambda: ops.gemm(a, w, bias=bias, residual=res), lambda: ops.conv3x3(x8, w8)]
for _ in range(2):
    for f in fns: f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in fns: f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
