"""Frame-sharding parity: the frames of ONE clip over the ranks of this job, with NON-identity temporal layers
(temporal LoRA halo exchange, temporal-attention frames<->pixels exchange, K/V push, GroupNorm statistics exchange: fz_p2p.cu).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py [--big]
Checks (rank 0 prints one JSON line, exit code 1 on failure):
  * one CFG-batch UNet forward, sharded vs the same product on one GPU (same kernels: only the fp32 association of the GroupNorm sums and
    the pixel split of the temporal attention differ) — bound 4e-3 of max|eps|;
  * the committed GOLDEN cases of the unmodified reference whose frame count divides over the ranks, through the reference-facing
    pipeline API on the rank's frames — same bounds as the single-GPU tests (tests/test_gpu_pipeline.py, tests/test_gpu_golden_sd14.py);
  * CUDA-graph replay of the sharded loops (third run of a case) equals the eager sharded run bit for bit."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("FZ_P2P_ARENA_MB", "3072")
from fatezero_b200 import dist as fzdist, synth  # noqa: E402
from _helpers import GOLDEN_DIR, build_product, run_product_case  # noqa: E402
from oracle.cases import CASES  # noqa: E402


def gmax(v: float, dev) -> float:
    t = torch.tensor([float(v)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = fzdist.env_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    fzdist.init("nccl", dev)
    big = "--big" in sys.argv
    out, ok = dict(world=world), True
    # ---- (1) one forward, sharded vs single GPU, non-identity temporal layers, a relative index (all-gather path) + constant ones
    mc = dict(lora=160, SparseCausalAttention_index=["mid", -1, "first"], least_sc_channel=64)
    frames, size = 2 * world, 32
    pipe = build_product("mini", mc, device=dev)
    cfg = synth.UNET_CONFIGS["mini"]
    x0 = (synth.synth_latents(frames, size, size) * 0.5).to(dev)
    x2 = torch.cat([x0, 0.7 * x0])
    emb = torch.randn(2, 77, cfg["cross_attention_dim"], generator=torch.Generator().manual_seed(2)).to(dev)
    full = pipe.unet(x2, 481, emb).sample
    pipe.unet.set_frame_shard(rank, world)
    loc = pipe.unet(fzdist.frame_slice(x2, rank, world), 481, emb).sample
    got = fzdist.gather_frames(loc, world)
    pipe.unet.set_frame_shard(0, 1)
    out["forward_max_abs"] = (got - full).abs().max().item()
    out["forward_ref_max"] = full.abs().max().item()
    ok &= out["forward_max_abs"] <= 4e-3 * max(out["forward_ref_max"], 1.0)
    del pipe
    # ---- (2) golden cases of the unmodified reference on the rank's frames (+ (3) graph replay == eager)
    names = ["mini_replace_blend", "mini_reweight_next"] + (["sd14_replace_blend", "sd14_config1"] if big else [])
    bounds = dict(mini_replace_blend=(2e-2, 8e-2), mini_reweight_next=(2e-2, 8e-2), sd14_replace_blend=(9.2e-3, 4.7e-2), sd14_config1=(1.2e-2, 5.5e-1))
    for name in names:
        case = CASES[name]
        gpath = os.path.join(GOLDEN_DIR, f"{name}.pt")
        if case["frames"] % world or not os.path.exists(gpath):
            continue
        g = torch.load(gpath)
        blend = bool(case["p2p"].get("blend_words"))
        pipe = build_product(case["unet"], case["model_config"], device=dev)
        pipe.unet.set_frame_shard(rank, world)
        runs = [run_product_case(case, pipe=pipe, device=dev, shard=(rank, world)) for _ in range(3)]  # eager, captured, replayed
        g_inv = fzdist.frame_slice(g["inv_latents"], rank, world, dim=3)
        g_ed = fzdist.frame_slice(g["edit_latents"], rank, world, dim=3)
        p = runs[0]
        d_inv = (p["inv_latents"] - g_inv).abs().max().item()
        d = (p["edit_latents"][-1] - g_ed[-1]).abs().reshape(-1)
        d_ed = (torch.quantile(d, 0.99) if blend else d.max()).item()
        sharded_abs = name.startswith("sd14")
        s_inv = 1.0 if sharded_abs else g["inv_latents"].abs().max().item()
        s_ed = 1.0 if sharded_abs else g["edit_latents"][-1].abs().max().item()
        e_inv, e_ed = gmax(d_inv / s_inv, dev), gmax(d_ed / s_ed, dev)
        same = all(torch.equal(r["inv_latents"], p["inv_latents"]) and torch.equal(r["edit_latents"], p["edit_latents"]) for r in runs[1:])
        same = gmax(0.0 if same else 1.0, dev) == 0.0
        out[name] = dict(inv_err=e_inv, edit_err=e_ed, graph_replay_equal=same, plans=len(pipe._plans))
        ok &= e_inv <= bounds[name][0] and e_ed <= bounds[name][1] and same and len(pipe._plans) == 2
        del pipe, runs
        torch.cuda.empty_cache()
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
