"""Run the UNMODIFIED FateZero reference (/root/reference) on CPU through the tests-only shim.  Build container only.

Follows the harness recipe of SURVEY.md App. C ("Harness recipe used"); cites:
  pipelines/p2p_ddim_spatial_temporal.py:68-148 (inversion), :172-222,:260-435 (edit),
  prompt_attention/attention_register.py:14 (hook registration).
Nothing here travels to the GPU box (it needs /root/reference); outputs are committed as tests/golden/*.pt.
"""
from __future__ import annotations

import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("FZ_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "video_diffusion"))


def _prepare_imports():
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    # the repo ships a drop-in `video_diffusion` alias package: make sure the REFERENCE one wins in this process
    for k in [k for k in sys.modules if k == "video_diffusion" or k.startswith("video_diffusion.")]:
        del sys.modules[k]
    for p in (_SHIM, REFERENCE_ROOT):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # a regular package (the repo's alias) beats the reference's namespace package in the path search: pin the name explicitly
    import types
    pkg = types.ModuleType("video_diffusion")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "video_diffusion")]
    sys.modules["video_diffusion"] = pkg
    # models/attention.py:229 calls torch.cuda.get_device_name(0) for every block
    if not torch.cuda.is_available():
        torch.cuda.get_device_name = lambda *a, **k: "cpu-shim"


def build_reference_pipeline(unet_config: dict, model_config: dict, seed: int = 0, degenerate_temporal: bool = False,
                             text_dim: int | None = None):
    """Reference UNet (+pipeline) with name-keyed synthetic weights (fatezero_b200.synth recipe)."""
    _prepare_imports()
    from fatezero_b200 import synth
    from video_diffusion.models.unet_3d_condition import UNetPseudo3DConditionModel
    from video_diffusion.pipelines.p2p_ddim_spatial_temporal import P2pDDIMSpatioTemporalPipeline
    from diffusers.schedulers import DDIMScheduler

    cfg = dict(unet_config)
    cfg.update(model_config or {})
    unet = UNetPseudo3DConditionModel(**cfg)
    shapes = {k: tuple(v.shape) for k, v in unet.state_dict().items()}
    unet.load_state_dict(synth.synth_state_dict(shapes, seed, degenerate_temporal))
    unet.eval().requires_grad_(False)
    text_dim = text_dim or unet_config["cross_attention_dim"]
    sched = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                          clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    pipe = P2pDDIMSpatioTemporalPipeline(vae=synth.VaeStub(), text_encoder=synth.ToyTextEncoder(text_dim, seed),
                                         tokenizer=synth.ToyTokenizer(), unet=unet, scheduler=sched)
    return pipe


@torch.no_grad()
def run_reference(pipe, x0, source_prompt: str, target_prompt: str, num_steps: int, p2p_config: dict,
                  guidance_scale: float = 7.5, save_path: str | None = None, capture_maps: bool = True):
    """Inversion (store on, LOW_RESOURCE) then edit_type='swap'. Returns dict of tensors (all CPU fp32)."""
    from video_diffusion.prompt_attention import attention_util
    pipe.scheduler.set_timesteps(num_steps)
    emb = pipe._encode_prompt(source_prompt, torch.device("cpu"), 1, True, None)
    pipe.prepare_before_train_loop()
    pipe.store_controller = attention_util.AttentionStore()
    attention_util.register_attention_control(pipe, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = True
    inv = pipe.ddim_clean2noisy_loop(x0, emb, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = False
    out = {"inv_latents": torch.stack([l.float() for l in inv], 0)}
    if capture_maps:
        maps = {}
        for step, d in enumerate(pipe.store_controller.attention_store_all_step):
            for key, lst in d.items():
                for pos, t in enumerate(lst):
                    maps[f"{step}/{key}/{pos}"] = t.float()
        out["maps"] = maps
    trace = []
    F = x0.shape[2]
    h, w = x0.shape[-2:]
    res = pipe(prompt=target_prompt, source_prompt=source_prompt, edit_type="swap", image=None, strength=None,
               generator=None, num_inference_steps=num_steps, clip_length=F, guidance_scale=guidance_scale,
               num_images_per_prompt=1, latents=inv[-1], uncond_embeddings_list=None, save_path=save_path,
               height=8 * h, width=8 * w, output_type="np",
               callback=lambda i, t, l: trace.append(l.detach().float().clone()),
               use_inversion_attention=True, save_self_attention=False, **p2p_config)
    out["edit_latents"] = torch.stack(trace, 0)
    out["mask_list"] = res["mask_list"]
    return out


@torch.no_grad()
def reference_unet_forward(pipe, x, t, emb):
    """One plain UNet forward with the dummy controller (attention_register.py:222-230)."""
    from video_diffusion.prompt_attention import attention_util
    attention_util.register_attention_control(pipe, None)
    return pipe.unet(x, t, encoder_hidden_states=emb).sample
