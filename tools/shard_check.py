"""Frame-sharding parity: the frames of one clip over the ranks of this job versus the same clip on one GPU.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py
Checks (rank 0 prints one JSON line, exit code 1 on failure):
  * one CFG-batch UNet forward: eps of the sharded run == eps of the single-GPU run up to the fp32 re-association of the GroupNorm sums
  * inversion + attention-fused edit of a mini case through the reference-facing pipeline API on the rank's frames
    (bounds: forward 4e-3, inversion 5e-3, edit 5e-2 relative; measured 1.5e-3 / 1.9e-3 / 2.1e-2 — two fp16 runs whose GroupNorm sums are associated differently; the CFG x7.5
    sampler amplifies that like any fp16 perturbation, cf. the 1.6e-2 of the single-GPU run against the fp32 oracle)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fatezero_b200 import dist as fzdist, synth  # noqa: E402
from _helpers import build_product  # noqa: E402


def main():
    rank, world, local = fzdist.env_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    fzdist.init("nccl", dev)
    # FZ_SHARD_INDEX=const: only constant source frames ('mid', 'first') -> the broadcast path; default: a relative index too -> all-gather
    index = ["mid", "first"] if os.environ.get("FZ_SHARD_INDEX") == "const" else ["mid", -1, "first"]
    mc = dict(lora=160, SparseCausalAttention_index=index, least_sc_channel=64)
    frames, size = 4, 32
    pipe = build_product("mini", mc, device=dev, degenerate_temporal=True)
    cfg = synth.UNET_CONFIGS["mini"]
    x0 = (synth.synth_latents(frames, size, size) * 0.5).to(dev)
    x2 = torch.cat([x0, 0.7 * x0])
    emb = torch.randn(2, 77, cfg["cross_attention_dim"], generator=torch.Generator().manual_seed(2)).to(dev)
    out = {}
    # ---- single-GPU reference on every rank
    full = pipe.unet(x2, 481, emb).sample
    # ---- sharded
    pipe.unet.set_frame_shard(rank, world)
    loc = pipe.unet(fzdist.frame_slice(x2, rank, world), 481, emb).sample
    got = fzdist.gather_frames(loc, world)
    pipe.unet.set_frame_shard(0, 1)
    out["forward_max_abs"] = (got - full).abs().max().item()
    out["forward_ref_max"] = full.abs().max().item()
    # ---- pipeline: inversion + edit on the local frames
    from fatezero_b200 import controllers
    src, tgt = "a silver jeep driving down a curvy road", "a watercolor painting of a silver jeep driving down a curvy road"
    p2p = dict(cross_replace_steps={"default_": 0.8}, self_replace_steps=0.7, is_replace_controller=False,
               eq_params=dict(words=["watercolor"], values=[4.0]), blend_words=None, use_inversion_attention=True)
    N = 4

    def run(x):
        pipe.scheduler.set_timesteps(N)
        e = pipe._encode_prompt(src, dev, 1, True, None)
        store = controllers.AttentionStore()
        pipe.store_controller = store
        controllers.register_attention_control(pipe, store)
        store.LOW_RESOURCE = True
        inv = pipe.ddim_clean2noisy_loop(x, e, store)
        store.LOW_RESOURCE = False
        o = pipe(prompt=tgt, source_prompt=src, edit_type="swap", image=None, strength=None, generator=None, num_inference_steps=N,
                 clip_length=x.shape[2], guidance_scale=7.5, num_images_per_prompt=1, latents=inv[-1], uncond_embeddings_list=None,
                 save_path=None, height=8 * size, width=8 * size, output_type="latent", save_self_attention=False, **p2p)
        return inv[-1], o["sdimage_output"].images

    inv_full, edit_full = run(x0)
    pipe.unet.set_frame_shard(rank, world)
    inv_loc, edit_loc = run(fzdist.frame_slice(x0, rank, world))
    inv_got, edit_got = fzdist.gather_frames(inv_loc, world), fzdist.gather_frames(edit_loc, world)
    pipe.unet.set_frame_shard(0, 1)
    out["inv_rel"] = ((inv_got - inv_full).abs().max() / inv_full.abs().max()).item()
    out["edit_rel"] = ((edit_got - edit_full).abs().max() / edit_full.abs().max()).item()
    ok = out["forward_max_abs"] <= 4e-3 * max(out["forward_ref_max"], 1.0) and out["inv_rel"] <= 5e-3 and out["edit_rel"] <= 5e-2
    out.update(world=world, frames=frames, index=index, ok=bool(ok))
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
