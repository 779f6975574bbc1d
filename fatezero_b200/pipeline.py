"""P2pDDIMSpatioTemporalPipeline — drop-in for video_diffusion/pipelines/p2p_ddim_spatial_temporal.py (+ the parts of
pipelines/stable_diffusion.py it inherits), same constructor, methods, kwargs and return types, so test_fatezero.py and
P2pSampleLogger (pipelines/p2p_validation_loop.py) drive it unchanged.

The two hot loops run on the GPU through libfatezero_b200.so:
  ddim_clean2noisy_loop  : N x { UNet (B=1, STORE fused in attention) ; fz_ddim_invert_step }
  sd_ddim_pipeline loop  : N x { UNet (B=2, INJECT/BLEND fused)       ; fz_cfg_ddim_step (+ latent blend) }
Latents stay fp32 in HBM for the whole run; text encoder / VAE / tokenizer are whatever objects the caller passes (out of the
accelerated scope).
"""
from __future__ import annotations

import inspect
import os
import sys
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from . import controllers as attention_util
from . import ops
from .scheduler import DDIMScheduler


class StableDiffusionPipelineOutput(dict):
    def __init__(self, images, nsfw_content_detected=None):
        super().__init__(images=images, nsfw_content_detected=nsfw_content_detected)
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected


class _NullBar:
    def __init__(self, total=None):
        self.total = total

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, n=1):
        pass


class SpatioTemporalStableDiffusionPipeline:
    """pipelines/stable_diffusion.py:33-336 (the members the FateZero flow uses)."""
    _optional_components: List[str] = []

    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler):
        cfg = getattr(scheduler, "config", None)
        # scheduler-config fix-ups of stable_diffusion.py:56-81
        if cfg is not None and getattr(cfg, "steps_offset", 1) != 1:
            new = dict(cfg)
            new["steps_offset"] = 1
            scheduler._internal_dict = type(cfg)(new)
        if cfg is not None and getattr(cfg, "clip_sample", False) is True:
            new = dict(scheduler.config)
            new["clip_sample"] = False
            scheduler._internal_dict = type(scheduler.config)(new)
        self.vae, self.text_encoder, self.tokenizer, self.unet, self.scheduler = vae, text_encoder, tokenizer, unet, scheduler
        self._config = {k: (type(v).__module__, type(v).__name__) for k, v in
                        dict(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, unet=unet, scheduler=scheduler).items()}
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self._progress_bar_config = {}

    # ---- DiffusionPipeline surface ------------------------------------------------------------------------------
    @property
    def config(self):
        return self._config

    @property
    def device(self):
        for m in (self.unet, self.text_encoder, self.vae):
            if isinstance(m, torch.nn.Module):
                try:
                    return next(m.parameters()).device
                except StopIteration:
                    continue
        return torch.device("cpu")

    @property
    def _execution_device(self):
        return self.device

    def to(self, device):
        for m in (self.unet, self.text_encoder, self.vae):
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    def progress_bar(self, iterable=None, total=None):
        return _NullBar(total)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None  # the fused sm_100a attention kernel is always on

    def disable_xformers_memory_efficient_attention(self, *a, **k):
        return None

    def enable_vae_slicing(self):
        if hasattr(self.vae, "enable_slicing"):
            self.vae.enable_slicing()

    def disable_vae_slicing(self):
        if hasattr(self.vae, "disable_slicing"):
            self.vae.disable_slicing()

    @staticmethod
    def numpy_to_pil(images):
        """stable_diffusion.py:566-576: [b, f, h, w, c] -> list (per clip) of lists of PIL frames."""
        from PIL import Image

        def frames(arr):
            if arr.ndim == 3:
                arr = arr[None, ...]
            arr = (arr * 255).round().astype("uint8")
            return [Image.fromarray(a) for a in arr]
        if images.ndim == 5:
            return [frames(seq) for seq in images]
        return [frames(images)]

    @staticmethod
    def _get_signature_keys(obj):
        params = inspect.signature(obj.__init__).parameters
        required = {k for k, v in params.items() if v.default is inspect._empty} - {"self"}
        optional = {k for k, v in params.items() if v.default is not inspect._empty}
        return required, optional

    def prepare_before_train_loop(self, params_to_optimize=None):
        for m in (self.vae, self.unet, self.text_encoder):
            if isinstance(m, torch.nn.Module):
                m.requires_grad_(False)
                m.eval()
        if params_to_optimize is not None:
            params_to_optimize.requires_grad = True

    def _text_forward(self, ids, mask):
        """The text encoder call of stable_diffusion.py:230,279.  A transformers CLIPTextModel living on the GPU is executed by
        clip.ClipTextEngine (sm_100a kernels, built once per module); any other module — or a padding mask — is simply called."""
        te = self.text_encoder
        if (mask is None and os.environ.get("FZ_CLIP", "1") != "0" and type(te).__name__ == "CLIPTextModel" and isinstance(te, torch.nn.Module)
                and next(te.parameters()).is_cuda):
            eng = getattr(self, "_clip_engine", None)
            if eng is None or eng[0] is not te:
                from .clip import ClipTextEngine
                try:
                    eng = (te, ClipTextEngine(te))
                except NotImplementedError:
                    eng = (te, None)
                self._clip_engine = eng
            if eng[1] is not None:
                return eng[1](ids)[0]
        return te(ids, attention_mask=mask)[0]

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt):
        """stable_diffusion.py:180-295: [uncond ; cond] text embeddings of shape [2*b, 77, D]."""
        batch_size = len(prompt) if isinstance(prompt, list) else 1
        tok = self.tokenizer
        text_inputs = tok(prompt, padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt")
        ids = text_inputs.input_ids
        use_mask = bool(getattr(getattr(self.text_encoder, "config", None), "use_attention_mask", False))
        mask = text_inputs.attention_mask.to(device) if use_mask else None
        emb = self._text_forward(ids.to(device), mask)
        bs, seq, _ = emb.shape
        emb = emb.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_classifier_free_guidance:
            if negative_prompt is None:
                uncond_tokens = [""] * batch_size
            elif type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} != {type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond_tokens = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but `prompt`: {prompt} has "
                                 f"batch size {batch_size}. Please make sure that passed `negative_prompt` matches the batch size of `prompt`.")
            else:
                uncond_tokens = negative_prompt
            un = tok(uncond_tokens, padding="max_length", max_length=ids.shape[-1], truncation=True, return_tensors="pt")
            umask = un.attention_mask.to(device) if use_mask else None
            uemb = self._text_forward(un.input_ids.to(device), umask)
            uemb = uemb.repeat(1, num_images_per_prompt, 1).view(batch_size * num_images_per_prompt, uemb.shape[1], -1)
            emb = torch.cat([uemb, emb])
        return emb

    def _vae_engine(self):
        """The sm_100a VAE executor for `self.vae` (fatezero_b200.vae): our AutoencoderKL container or a foreign AutoencoderKL-shaped
        module on the GPU; None for anything else (stubs, CPU modules) — those are simply called."""
        if os.environ.get("FZ_VAE", "1") == "0":
            return None
        cached = getattr(self, "_vae_cache", None)
        if cached is None or cached[0] is not self.vae:
            from . import vae as vae_mod
            cached = (self.vae, vae_mod.engine_for(self.vae))
            self._vae_cache = cached
        return cached[1]

    def _vae_encode_sample(self, image, generator):
        """p2p_ddim_spatial_temporal.py:88-96: vae.encode(image).latent_dist.sample(generator)."""
        eng = self._vae_engine()
        if eng is None:
            return self.vae.encode(image).latent_dist.sample(generator)
        from .vae import DiagonalGaussianDistribution
        return DiagonalGaussianDistribution(eng.encode_moments(image).to(image.dtype)).sample(generator)

    def decode_latents(self, latents):
        """stable_diffusion.py:297-319 (VAE decode in chunks of 16 frames)."""
        is_video = latents.dim() == 5
        b = latents.shape[0]
        latents = 1 / 0.18215 * latents
        if is_video:
            latents = latents.permute(0, 2, 1, 3, 4).reshape(-1, *latents.shape[1:2], *latents.shape[3:])
        vdt = next(self.vae.parameters()).dtype if isinstance(self.vae, torch.nn.Module) else latents.dtype
        eng = self._vae_engine()
        if eng is not None:
            image = torch.cat([eng.decode(chunk) for chunk in torch.split(latents, 16, dim=0)], dim=0)
        else:
            image = torch.cat([self.vae.decode(chunk.to(vdt)).sample for chunk in torch.split(latents, 16, dim=0)], dim=0)
        image = (image / 2 + 0.5).clamp(0, 1).cpu().float().numpy()
        if is_video:
            image = image.reshape(b, -1, *image.shape[1:]).transpose(0, 1, 3, 4, 2)
        else:
            image = image.transpose(0, 2, 3, 1)
        return image

    def prepare_extra_step_kwargs(self, generator, eta):
        keys = set(inspect.signature(self.scheduler.step).parameters.keys())
        out = {}
        if "eta" in keys:
            out["eta"] = eta
        if "generator" in keys:
            out["generator"] = generator
        return out


class P2pDDIMSpatioTemporalPipeline(SpatioTemporalStableDiffusionPipeline):
    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler, disk_store: bool = False):
        super().__init__(vae, text_encoder, tokenizer, unet, scheduler)
        self.store_controller = attention_util.AttentionStore(disk_store=disk_store)
        self.empty_controller = attention_util.EmptyControl()
        # CUDA-graph execution of the two DDIM loops (graphs.py): "auto" = eager the first time a configuration is seen, captured and
        # replayed from its second occurrence on; "off" = always eager
        self.graph_mode = os.environ.get("FZ_GRAPHS", "auto")
        self._plans = {}
        self._seen = set()

    def release_graphs(self):
        """Drop every captured plan (and the HBM pools that hold their map caches)."""
        self._plans.clear()
        self._seen.clear()

    def check_inputs(self, prompt, height, width, callback_steps, strength=None):
        if not isinstance(prompt, str) and not isinstance(prompt, list):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if strength is not None and (strength <= 0 or strength > 1):
            raise ValueError(f"The value of strength should in (0.0, 1.0] but is {strength}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if (callback_steps is None) or (callback_steps is not None and (not isinstance(callback_steps, int) or callback_steps <= 0)):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type {type(callback_steps)}.")

    # ---- scheduler tables (work with diffusers' DDIMScheduler or fatezero_b200.scheduler.DDIMScheduler) -------------
    def _alpha(self, t: int) -> float:
        return float(self.scheduler.alphas_cumprod[int(t)]) if int(t) >= 0 else float(self.scheduler.final_alpha_cumprod)

    # ---- inversion ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prepare_latents_ddim_inverted(self, image, batch_size, num_images_per_prompt, text_embeddings, store_attention=False,
                                      prompt=None, generator=None, LOW_RESOURCE=True, save_path=None):
        """p2p_ddim_spatial_temporal.py:68-129."""
        self.prepare_before_train_loop()
        if store_attention:
            attention_util.register_attention_control(self, self.store_controller)
        resource_default_value = self.store_controller.LOW_RESOURCE
        self.store_controller.LOW_RESOURCE = LOW_RESOURCE
        batch_size = batch_size * num_images_per_prompt
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch size of "
                             f"{batch_size}. Make sure the batch size matches the length of the generators.")
        if isinstance(generator, list):
            init_latents = torch.cat([self._vae_encode_sample(image[i:i + 1], generator[i]) for i in range(batch_size)], dim=0)
        else:
            init_latents = self._vae_encode_sample(image, generator)
        init_latents = 0.18215 * init_latents
        if batch_size > init_latents.shape[0] and batch_size % init_latents.shape[0] != 0:
            raise ValueError(f"Cannot duplicate `image` of batch size {init_latents.shape[0]} to {batch_size} text prompts.")
        if batch_size > init_latents.shape[0]:
            init_latents = torch.cat([init_latents] * (batch_size // init_latents.shape[0]), dim=0)
        bf, c, h, w = init_latents.shape
        init_bcfhw = init_latents.reshape(batch_size, bf // batch_size, c, h, w).permute(0, 2, 1, 3, 4)
        out = self.ddim_clean2noisy_loop(init_bcfhw, text_embeddings, self.store_controller)
        if store_attention and (save_path is not None):
            os.makedirs(save_path + "/cross_attention")
            from .visualization import show_cross_attention
            show_cross_attention(self.tokenizer, prompt, self.store_controller, 16, ["up", "down"], save_path=save_path + "/cross_attention")
            attention_util.register_attention_control(self, self.empty_controller)
        self.store_controller.LOW_RESOURCE = resource_default_value
        return out

    @torch.no_grad()
    def ddim_clean2noisy_loop(self, latent, text_embeddings, controller=None, teacher_latents=None):
        """p2p_ddim_spatial_temporal.py:131-148.  Returns N+1 latents (dtype of the input), [0] clean, [-1] x_T.
        teacher_latents (parity tests only): N+1 reference latents; step i then starts from teacher_latents[i] instead of this loop's
        own x_i (and the controller stores teacher_latents[i+1]), which isolates the per-forward kernel error from its amplification
        by the sampler."""
        weight_dtype = latent.dtype
        dev = self.unet.device
        ts = [int(t) for t in self.scheduler.timesteps]
        n = len(ts)
        step = self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps
        with torch.cuda.device(dev):
            cond = text_embeddings.chunk(2)[1].to(dev).contiguous()

            def inv_step(i, x, cond_, ctrl, outputs):
                t = ts[n - 1 - i]
                eps = self.unet(x, t, encoder_hidden_states=cond_)["sample"]
                ops.ddim_invert_step(x, eps.contiguous(), self._alpha(min(t - step, 999)), self._alpha(t))
                if ctrl is not None:
                    ctrl.step_callback(x if teacher_latents is None else teacher_latents[i + 1].to(dev, torch.float32))
                outputs.append(x.to(dtype=weight_dtype).clone())

            # ---- CUDA-graph path (graphs.py): same launch sequence, captured once per configuration, no Python in the step ----
            sig = None
            if (self.graph_mode != "off" and teacher_latents is None and isinstance(controller, attention_util.AttentionStore)
                    and getattr(self.unet, "_controller", None) is controller and controller.is_pristine()):
                sig = controller.graph_signature()
            key = None if sig is None else ("inv", tuple(latent.shape), str(weight_dtype), tuple(ts), tuple(cond.shape), sig,
                                            id(self.unet.engine()), self.unet.engine().shard_signature())
            if key is not None and (key in self._plans or key in self._seen):
                plan = self._plans.get(key)
                if plan is None:
                    from .graphs import LoopPlan
                    plan = LoopPlan(dev)
                    plan.x = torch.empty(latent.shape, dtype=torch.float32, device=dev)
                    plan.text = torch.empty_like(cond)
                    plan.controller = controller
                    plan.x.copy_(latent)
                    plan.text.copy_(cond)
                    for i in range(n):
                        plan.steps.capture(lambda i=i: inv_step(i, plan.x, plan.text, controller, plan.outputs))
                    controller._graph_plan_id = plan.id
                    # the capture advanced the controller's Python state to the end of the loop; the replay below fills its tensors
                    self._plans[key] = plan
                plan.x.copy_(latent)
                plan.text.copy_(cond)
                for i in range(n):
                    plan.steps.replay(i)
                controller.adopt_from(plan.controller)
                return [latent] + [o.clone() for o in plan.outputs]
            if key is not None:
                self._seen.add(key)
            all_latent = [latent]
            x = latent.detach().to(dev, torch.float32).contiguous().clone()
            for i in range(n):
                if teacher_latents is not None:
                    x.copy_(teacher_latents[i].to(dev, torch.float32))
                inv_step(i, x, cond, controller, all_latent)
        return all_latent

    def next_clean2noise_step(self, model_output, timestep, sample):
        """p2p_ddim_spatial_temporal.py:150-161 on tensors (the loop above uses the fused kernel with the same coefficients)."""
        step = self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps
        timestep, next_timestep = min(int(timestep) - step, 999), int(timestep)
        x = sample.detach().to(torch.float32).contiguous().clone()
        ops.ddim_invert_step(x, model_output.to(torch.float32).contiguous(), self._alpha(timestep), self._alpha(next_timestep))
        return x

    def get_timesteps(self, num_inference_steps, strength, device):
        init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
        t_start = max(num_inference_steps - init_timestep, 0)
        return self.scheduler.timesteps[t_start:], num_inference_steps - t_start

    # ---- edit -----------------------------------------------------------------------------------------------------------
    def p2preplace_edit(self, **kwargs):
        """p2p_ddim_spatial_temporal.py:172-222."""
        len_source = len(kwargs["source_prompt"].split(" "))
        len_target = len(kwargs["prompt"].split(" "))
        equal_length = len_source == len_target
        edit_controller = attention_util.make_controller(
            self.tokenizer, [kwargs["source_prompt"], kwargs["prompt"]], NUM_DDIM_STEPS=kwargs["num_inference_steps"],
            is_replace_controller=kwargs.get("is_replace_controller", True) and equal_length,
            cross_replace_steps=kwargs["cross_replace_steps"], self_replace_steps=kwargs["self_replace_steps"],
            blend_words=kwargs.get("blend_words", None), equilizer_params=kwargs.get("eq_params", None),
            additional_attention_store=self.store_controller, use_inversion_attention=kwargs["use_inversion_attention"],
            blend_th=kwargs.get("blend_th", (0.3, 0.3)), blend_self_attention=kwargs.get("blend_self_attention", None),
            blend_latents=kwargs.get("blend_latents", None), save_path=kwargs.get("save_path", None),
            save_self_attention=kwargs.get("save_self_attention", True), disk_store=kwargs.get("disk_store", False))
        attention_util.register_attention_control(self, edit_controller)
        sdimage_output = self.sd_ddim_pipeline(controller=edit_controller, **kwargs)
        mask_list = edit_controller.latent_blend.mask_list if hasattr(edit_controller.latent_blend, "mask_list") else None
        attention_output = None
        if len(edit_controller.attention_store.keys()) > 0 and kwargs.get("output_type", "pil") != "latent":
            from .visualization import show_cross_attention
            attention_output = show_cross_attention(self.tokenizer, kwargs["prompt"], edit_controller, 16, ["up", "down"])
        self.last_edit_controller = edit_controller
        attention_util.register_attention_control(self, self.empty_controller)
        return {"sdimage_output": sdimage_output, "attention_output": attention_output, "mask_list": mask_list}

    @torch.no_grad()
    def __call__(self, **kwargs):
        edit_type = kwargs["edit_type"]
        assert edit_type in ["save", "swap", None]
        if edit_type is None:
            return self.sd_ddim_pipeline(controller=None, **kwargs)
        if edit_type == "save":
            self.store_controller = attention_util.AttentionStore()
            attention_util.register_attention_control(self, self.store_controller)
            sdimage_output = self.sd_ddim_pipeline(controller=self.store_controller, **kwargs)
            from .visualization import show_cross_attention
            attention_output = show_cross_attention(self.tokenizer, kwargs["prompt"], self.store_controller, 16, ["up", "down"])
            attention_util.register_attention_control(self, self.empty_controller)
            return {"sdimage_output": sdimage_output, "attention_output": attention_output, "mask_list": None}
        return self.p2preplace_edit(**kwargs)

    @torch.no_grad()
    def sd_ddim_pipeline(self, prompt: Union[str, List[str]], image=None, height: Optional[int] = None, width: Optional[int] = None,
                         strength: float = None, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                         negative_prompt: Optional[Union[str, List[str]]] = None, num_images_per_prompt: Optional[int] = 1,
                         eta: float = 0.0, generator=None, latents: Optional[torch.FloatTensor] = None, output_type: Optional[str] = "pil",
                         return_dict: bool = True, callback: Optional[Callable[[int, int, torch.FloatTensor], None]] = None,
                         callback_steps: Optional[int] = 1, controller=None, teacher_latents=None, **args):
        """p2p_ddim_spatial_temporal.py:260-435 (unknown kwargs are swallowed like the reference's **args).
        teacher_latents (parity tests only): reference latents AFTER each step; step i > 0 then starts from teacher_latents[i - 1]."""
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps, strength)
        if eta != 0.0:
            raise NotImplementedError("FateZero's DDIM path is deterministic (eta = 0)")
        batch_size = 1 if isinstance(prompt, str) else len(prompt)
        device = self._execution_device
        do_cfg = guidance_scale > 1.0
        if not do_cfg:
            raise NotImplementedError("guidance_scale <= 1 (no CFG batch) is not an editing configuration of the reference YAMLs")
        text_embeddings = self._encode_prompt(prompt, device, num_images_per_prompt, do_cfg, negative_prompt).to(self.unet.device)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = [int(t) for t in self.scheduler.timesteps]
        if latents is None:
            # the internal inversion runs batch 1 through the UNet: the edit controller registered by p2preplace_edit must not see it
            # (the reference's inversion hooks run the store logic of the edit controller on that pass; here it is detached)
            registered = getattr(self.unet, "_controller", None)
            attention_util.register_attention_control(self, self.empty_controller)
            try:
                latents = self.prepare_latents_ddim_inverted(image, batch_size, num_images_per_prompt, text_embeddings,
                                                             store_attention=False, generator=generator)[-1]
            finally:
                attention_util.register_attention_control(self, registered)
        latents_dtype = latents.dtype
        dev = self.unet.device
        text_embeddings = text_embeddings.contiguous()
        is_edit = isinstance(controller, attention_util.AttentionControlEdit)
        step = self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps
        n = len(timesteps)

        def edit_step(i, x, text, ctrl):
            t = timesteps[i]
            x2 = torch.cat([x, x], dim=0)
            eps2 = self.unet(x2, t, encoder_hidden_states=text).sample
            blend = ctrl.latent_blend_args(x.shape[-2], x.shape[-1]) if is_edit else None
            a_t, a_prev = self._alpha(t), self._alpha(t - step)
            if blend is not None:
                ops.cfg_ddim_step(x, eps2.contiguous(), guidance_scale, a_t, a_prev, x_inv=blend["x_inv"].contiguous(),
                                  mask_a=blend["mask_a"], mask_b=blend["mask_b"], apply_blend=blend["apply_blend"])
            else:
                ops.cfg_ddim_step(x, eps2.contiguous(), guidance_scale, a_t, a_prev)
            if ctrl is not None:
                if is_edit:
                    ctrl.step_callback(x, blend_fused=True)
                else:
                    ctrl.step_callback(x)

        with torch.cuda.device(dev):
            sig = None
            if (self.graph_mode != "off" and teacher_latents is None and is_edit and getattr(self.unet, "_controller", None) is controller
                    and controller.cur_step == 0):
                sig = controller.graph_signature()
            key = None if sig is None else ("edit", tuple(latents.shape), tuple(timesteps), tuple(text_embeddings.shape), float(guidance_scale),
                                            sig[:-1], id(self.unet.engine()), self.unet.engine().shard_signature())
            store_plan = None if sig is None else sig[-1]
            if key is not None and store_plan is not None and ((key, store_plan) in self._plans or key in self._seen):
                plan = self._plans.get((key, store_plan))
                if plan is None:
                    from .graphs import LoopPlan
                    plan = LoopPlan(dev)
                    plan.x = torch.empty(latents.shape, dtype=torch.float32, device=dev)
                    plan.text = torch.empty_like(text_embeddings)
                    plan.controller = controller
                    controller.prepare_tables(dev)
                    plan.x.copy_(latents)
                    plan.text.copy_(text_embeddings)
                    for i in range(n):
                        plan.steps.capture(lambda i=i: edit_step(i, plan.x, plan.text, controller))
                    self._plans[(key, store_plan)] = plan
                plan.x.copy_(latents)
                plan.text.copy_(text_embeddings)
                plan.controller.load_tables_from(controller)
                for i in range(n):
                    plan.steps.replay(i)
                    if callback is not None and i % callback_steps == 0:
                        callback(i, timesteps[i], plan.x.to(latents_dtype))
                controller.adopt_from(plan.controller)
                x = plan.x.clone()
            else:
                if key is not None:
                    self._seen.add(key)
                x = latents.detach().to(dev, torch.float32).contiguous().clone()
                for i, t in enumerate(timesteps):
                    if teacher_latents is not None and i > 0:
                        x.copy_(teacher_latents[i - 1].to(x.device, torch.float32))
                    edit_step(i, x, text_embeddings, controller)
                    if callback is not None and i % callback_steps == 0:
                        callback(i, t, x.to(latents_dtype))
        latents = x.to(latents_dtype)
        if output_type == "latent":
            return StableDiffusionPipelineOutput(images=latents, nsfw_content_detected=None)
        image = self.decode_latents(latents)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
        if not return_dict:
            return (image, None)
        return StableDiffusionPipelineOutput(images=image, nsfw_content_detected=None)

    def print_pipeline(self, logger):
        print("Overview function of pipeline: ")
        print(self.__class__)
        logger.info(str({k: getattr(self, k).__class__ for k in self.config.keys()}))
        print(f"python version {sys.version}")
        print(f"torch version {torch.__version__}")
        print("validate gpu status:")
        print(torch.tensor(1.0).cuda() * 2)
        from . import _lib
        print(f"libfatezero_b200 version {_lib.load().fz_version()} at {_lib.LIB_PATH}")
