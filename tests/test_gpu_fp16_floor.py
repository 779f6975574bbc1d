"""How close to fp32 CAN an fp16 implementation of this forward be?  The reference runs the UNet under fp16 autocast on the GPU
(test_fatezero.py:153, mixed_precision "fp16"), so its own outputs deviate from the fp32 CPU run the goldens come from.  This test
measures that deviation with the oracle restatement executed by PyTorch on the GPU under torch.autocast(float16) (cuBLAS / cuDNN
kernels, fp32 softmax and norms exactly like the reference's autocast path) and compares the CUDA product's deviation with it: the
product must be no further from fp32 than 1.5x the reference's own fp16 mode.  (The oracle is the checker here, never the product.)"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from _helpers import GOLDEN_DIR, build_oracle, build_product  # noqa: E402
from fatezero_b200 import synth  # noqa: E402
from oracle.cases import CASES  # noqa: E402


def test_product_error_is_the_fp16_floor(report):
    name = "sd14_replace_blend"
    g = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"))
    case = CASES[name]
    x0 = synth.synth_latents(case["frames"], case["size"], case["size"]) * 0.5
    x2 = torch.cat([x0, 0.7 * x0])[:, :, :2].cuda()
    emb = torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(2)).cuda()
    ref32 = g["fwd_eps"]                                   # the UNMODIFIED reference, fp32, CPU
    ou = build_oracle(case["unet"], case["model_config"]).to("cuda")
    with torch.no_grad():
        o32 = ou.forward(x2, 481, emb).float().cpu()       # same restatement, fp32 on the GPU (TF32 off by default for matmul? conv may use it)
        with torch.autocast("cuda", dtype=torch.float16):
            o16 = ou.forward(x2, 481, emb).float().cpu()   # the reference's numerics on a GPU: fp16 autocast, fp32 weights
    del ou
    torch.cuda.empty_cache()
    pipe = build_product(case["unet"], case["model_config"])
    got = pipe.unet(x2, 481, emb).sample.float().cpu()
    d_gpu32 = (o32 - ref32).abs().max().item()
    d_auto = (o16 - ref32).abs().max().item()
    d_ours = (got - ref32).abs().max().item()
    rms = lambda a: (a - ref32).pow(2).mean().sqrt().item()  # noqa: E731
    report["fp16_floor"] = dict(torch_fp32_gpu_max_abs=d_gpu32, torch_fp16_autocast_max_abs=d_auto, product_max_abs=d_ours,
                                torch_fp16_autocast_rms=rms(o16), product_rms=rms(got), eps_abs_max=ref32.abs().max().item())
    print(f"\nmax|eps - fp32 reference|: torch fp32 on GPU {d_gpu32:.2e}, torch fp16 autocast {d_auto:.2e} (rms {rms(o16):.2e}), "
          f"fatezero_b200 {d_ours:.2e} (rms {rms(got):.2e})")
    assert d_ours <= 1.5 * d_auto + 1e-4, (d_ours, d_auto)
