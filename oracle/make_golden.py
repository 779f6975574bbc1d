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
    big = bool(c.get("big"))
    sq = {}
    for k, v in out["maps"].items():
        sums[k] = float(v.double().sum())
        step, key, pos = k.split("/")
        if big:
            # SD-1.4 geometry: 741 MiB of maps per step -> keep (a) the sum of squares of every map (sensitive to the distribution,
            # the plain sum of a softmax is just the row count), (b) slices of a few maps of the first and last step
            sq[k] = float((v.double() ** 2).sum())
            if step in ("0", str(c["steps"] - 1)):
                F = v.shape[0]
                if key.endswith("cross") and v.shape[2] <= 256:
                    keep[k] = v[F // 2].half().clone()                       # one frame, all heads   [8, s, 77]
                elif key.endswith("self") and v.shape[2] <= 256:
                    keep[k + "@f0h2"] = v[0, 2].half().clone()               # one (frame, head)      [s, t]
                elif key.endswith("self") and pos == "0":
                    keep[k + "@f1h5r256"] = v[min(1, F - 1), 5, :256].half().clone()  # r32: 256 query rows [256, 1024]
            continue
        if step == "0" and v.numel() <= 1 << 19 and len(keep) < (6 if c.get("gpu", True) else 2):  # pin-only cases: map checksums carry the rest
            keep[k] = v.half()
    gold["map_sums"] = sums
    if big:
        gold["map_sqsums"] = sq
    gold["maps"] = keep
    if out["mask_list"] is not None:
        gold["mask_list"] = [m.clone() for m in out["mask_list"]]
    # one plain forward for the single-forward parity tests
    emb = torch.randn(2, 77, cfg["cross_attention_dim"], generator=torch.Generator().manual_seed(2))
    x2 = torch.cat([x0, 0.7 * x0])
    if big:
        x2 = x2[:, :, :2]  # two frames are enough for the single-forward pin at SD-1.4 geometry
    gold["fwd_eps"] = rh.reference_unet_forward(pipe, x2, torch.tensor(481), emb).clone()
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(gold, path)
    print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", f"{gold['seconds']:.1f}s",
          "edit std", out["edit_latents"].std(dim=(1, 2, 3, 4, 5)).tolist(),
          "mask means", [float(m.mean()) for m in (out["mask_list"] or [])])


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
