"""Shared builders for the parity tests: the CUDA product, the CPU oracle and the case inputs (all from synthetic weights)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fatezero_b200 import synth  # noqa: E402
from fatezero_b200.unet import unet_param_spec  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def synth_weights(unet_name: str, model_config: dict, seed: int = 0, degenerate_temporal: bool = False):
    cfg = synth.UNET_CONFIGS[unet_name]
    spec = unet_param_spec(dict(cfg), model_config)
    return synth.synth_state_dict({k: v[0] for k, v in spec.items()}, seed, degenerate_temporal)


def build_oracle(unet_name: str, model_config: dict, **kw):
    from oracle import fz_oracle as fo
    return fo.OracleUNet(synth_weights(unet_name, model_config, **kw), synth.UNET_CONFIGS[unet_name], model_config)


def build_product(unet_name: str, model_config: dict, device="cuda", **kw):
    from fatezero_b200 import DDIMScheduler, P2pDDIMSpatioTemporalPipeline, UNetPseudo3DConditionModel
    cfg = synth.UNET_CONFIGS[unet_name]
    unet = UNetPseudo3DConditionModel(**cfg, **model_config)
    unet.load_state_dict(synth_weights(unet_name, model_config, **kw))
    unet.to(device)
    te = synth.ToyTextEncoder(cfg["cross_attention_dim"]).to(device)
    pipe = P2pDDIMSpatioTemporalPipeline(synth.VaeStub(), te, synth.ToyTokenizer(), unet, DDIMScheduler())
    return pipe


def case_inputs(case: dict):
    x0 = synth.synth_latents(case["frames"], case["size"], case["size"]) * 0.5
    return x0


def run_oracle_case(case: dict, ou=None):
    """Inversion + edit with the CPU oracle. Returns dict(inv_latents, edit_latents, store, ctrl)."""
    from oracle import fz_oracle as fo
    cfg = synth.UNET_CONFIGS[case["unet"]]
    ou = ou or build_oracle(case["unet"], case["model_config"])
    tok, te = synth.ToyTokenizer(), synth.ToyTextEncoder(cfg["cross_attention_dim"])
    emb_src = fo.encode_prompts(tok, te, case["source"])
    emb_tgt = fo.encode_prompts(tok, te, case["target"])
    x0 = case_inputs(case)
    N = case["steps"]
    store = fo.OracleStore()
    inv = fo.invert(ou, x0, emb_src[1:], N, store)
    p = case["p2p"]
    plan = fo.EditPlan(tok, case["source"], case["target"], N, p["cross_replace_steps"], p["self_replace_steps"],
                       p.get("is_replace_controller", True), p.get("eq_params"), p.get("blend_words"),
                       bool(p.get("blend_self_attention")), bool(p.get("blend_latents")), p.get("blend_th", (0.3, 0.3)))
    ctrl = fo.OracleEdit(plan, store)
    tr = fo.edit(ou, inv[-1], emb_tgt, N, ctrl)
    return dict(inv_latents=torch.stack(inv), edit_latents=torch.stack(tr), store=store, ctrl=ctrl)


def run_product_case(case: dict, pipe=None, device="cuda", save_path=None, teacher=None, shard=None):
    """Inversion + edit through the reference-facing API of the CUDA product.  teacher = golden dict: every forward starts from the
    reference's latent of that step (teacher forcing), so the returned latents carry ONE step of kernel error each."""
    import tempfile
    from fatezero_b200 import controllers
    pipe = pipe or build_product(case["unet"], case["model_config"], device)
    N = case["steps"]
    pipe.scheduler.set_timesteps(N)
    x0 = case_inputs(case).to(device)
    frames = case["frames"]
    if shard is not None:  # (rank, world): this process holds a contiguous block of the clip's frames (pipe.unet.set_frame_shard done by the caller)
        from fatezero_b200 import dist as fzdist
        x0 = fzdist.frame_slice(x0, shard[0], shard[1])
        frames = x0.shape[2]
    emb = pipe._encode_prompt(case["source"], device, 1, True, None)
    pipe.prepare_before_train_loop()
    pipe.store_controller = controllers.AttentionStore()
    controllers.register_attention_control(pipe, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = True
    inv = pipe.ddim_clean2noisy_loop(x0, emb, pipe.store_controller,
                                     teacher_latents=None if teacher is None else list(teacher["inv_latents"]))
    pipe.store_controller.LOW_RESOURCE = False
    trace = []
    p = dict(case["p2p"])
    if teacher is not None:
        p["teacher_latents"] = list(teacher["edit_latents"])
    if p.get("blend_words") and save_path is None:
        save_path = tempfile.mkdtemp()
    h = case["size"]
    res = pipe(prompt=case["target"], source_prompt=case["source"], edit_type="swap", image=None, strength=None, generator=None,
               num_inference_steps=N, clip_length=frames, guidance_scale=7.5, num_images_per_prompt=1,
               latents=inv[-1] if teacher is None else teacher["inv_latents"][-1].to(device),
               uncond_embeddings_list=None, save_path=save_path, height=8 * h, width=8 * h, output_type="latent",
               callback=lambda i, t, l: trace.append(l.detach().float().cpu().clone()), use_inversion_attention=True,
               save_self_attention=False, **p)
    return dict(inv_latents=torch.stack([l.float().cpu() for l in inv]), edit_latents=torch.stack(trace), pipe=pipe, result=res)
