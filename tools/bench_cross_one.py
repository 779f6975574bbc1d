"""Text cross-attention without a hook (77 keys) at the r=64 / r=32 shapes of the step: CUDA events, back-to-back launches."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fatezero_b200 import ops
dev = "cuda"
for (S, d, BF) in [(4096, 40, 16), (4096, 40, 8), (1024, 80, 16), (256, 160, 16)]:
    heads, B = 8, BF // 8
    C_ = heads * d
    q = torch.randn(BF * S, C_, device=dev).half(); k = torch.randn(B * 77, C_, device=dev).half()
    vt = torch.randn(B, heads, d, 80, device=dev).half(); out = torch.empty(BF * S, C_, device=dev, dtype=torch.float16)
    si = [[b for b in range(B) for _ in range(8)]]
    fn = lambda: ops.attention(q, k, vt, out, S_q=S, keys_per_slot=77, n_src=B, d=d, heads=heads, F=8, BF=BF, scale=d ** -0.5, src_index=si)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    print(dict(S=S, d=d, BF=BF, us=s.elapsed_time(e) / 20 * 1e3, variant=os.environ.get("FZ_LIB_VARIANT", "default")), flush=True)
