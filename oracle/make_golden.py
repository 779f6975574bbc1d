"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) through the tests-only shim.
Build container only (the reference does not travel).  Usage: python -m oracle.make_golden [case ...]"""
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fatezero_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle.cases import CASES  # noqa: E402


def run_case(name: str):
    c = CASES[name]
    cfg = synth.UNET_CONFIGS[c["unet"]]
    pipe = rh.build_reference_pipeline(cfg, c["model_config"])
    x0 = synth.synth_latents(c["frames"], c["size"], c["size"]) * 0.5
    save_path = tempfile.mkdtemp() if c["p2p"].get("blend_words") else None
    t = time.time()
    out = rh.run_reference(pipe, x0, c["source"], c["target"], c["steps"], c["p2p"], save_path=save_path)
    gold = dict(case=name, inv_latents=out["inv_latents"].clone(), edit_latents=out["edit_latents"].clone(),
                seconds=time.time() - t, torch=str(torch.__version__))
    # a few stored maps (fp16 to stay small) + checksums of all of them
    keep = {}
    sums = {}
    for k, v in out["maps"].items():
        sums[k] = float(v.double().sum())
        step, key, pos = k.split("/")
        if step == "0" and v.numel() <= 1 << 19 and len(keep) < (6 if c.get("gpu", True) else 2):  # pin-only cases: map checksums carry the rest
            keep[k] = v.half()
    gold["map_sums"] = sums
    gold["maps"] = keep
    if out["mask_list"] is not None:
        gold["mask_list"] = [m.clone() for m in out["mask_list"]]
    # one plain forward for the single-forward parity tests
    emb = torch.randn(2, 77, cfg["cross_attention_dim"], generator=torch.Generator().manual_seed(2))
    x2 = torch.cat([x0, 0.7 * x0])
    gold["fwd_eps"] = rh.reference_unet_forward(pipe, x2, torch.tensor(481), emb).clone()
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(gold, path)
    print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", f"{gold['seconds']:.1f}s",
          "edit std", out["edit_latents"].std(dim=(1, 2, 3, 4, 5)).tolist(),
          "mask means", [float(m.mean()) for m in (out["mask_list"] or [])])


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
