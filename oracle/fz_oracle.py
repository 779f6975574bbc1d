"""CPU fp32 restatement of the FateZero hot path (DDIM inversion + attention-fused denoising).  TEST INFRASTRUCTURE.

Portable (plain torch, device-agnostic, no /root/reference needed) so it travels to the GPU box as the parity checker for
the CUDA path.  It is *pinned* against the unmodified reference run through oracle/ref_harness.py (on a tests-only restatement of
diffusers 0.11.1, oracle/shim/): oracle/make_golden.py writes tests/golden/*.pt in the build container and
tests/test_oracle_golden.py holds this module to them.

Every function cites the reference lines it restates (paths relative to /root/reference/video_diffusion):
  UNet forward ............ models/unet_3d_condition.py:307-446, models/unet_3d_blocks.py:208,303,401,508,606
  resnet / pseudo-3D conv .. models/resnet.py:57-80,335-394 ; models/lora.py:46-54
  transformer .............. models/attention.py:95-144,271-337
  attention hooks .......... prompt_attention/attention_register.py:23-59,71-128,131-218
  store / edit controllers . prompt_attention/attention_store.py:38-49,81-109 ; attention_util.py:47-158,213-253,282-286
  blend masks .............. prompt_attention/spatial_blend.py:24-124
  host tables .............. prompt_attention/ptp_utils.py:144-199 ; seq_aligner.py:61-195 ; attention_util.py:307-316
  DDIM loops ............... pipelines/p2p_ddim_spatial_temporal.py:131-161,389-413 ; diffusers-0.11.1 DDIMScheduler
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# DDIM scheduler constants (diffusers 0.11.1 scheduling_ddim.py; SD-1.4 scheduler_config.json)
# --------------------------------------------------------------------------------------------------------------


def alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(n: int, num_train: int = 1000, steps_offset: int = 1) -> List[int]:
    ratio = num_train // n
    return [int(v) + steps_offset for v in (np.arange(0, n) * ratio).round()[::-1]]


# --------------------------------------------------------------------------------------------------------------
# UNet (functional, driven by a state dict with the reference's key names)
# --------------------------------------------------------------------------------------------------------------


def sc_frame_indices(index_list, clip_length: int) -> List[List[int]]:
    """attention_register.py:168-188: per K/V slot, the source frame of every query frame."""
    out = []
    for index in index_list:
        if isinstance(index, str):
            if index == "first":
                fi = [0] * clip_length
            elif index == "last":
                fi = [clip_length - 1] * clip_length
            elif index in ("mid", "middle"):
                fi = [int((clip_length - 1) // 2)] * clip_length
            else:
                raise ValueError(index)
        else:
            fi = [min(max(f + int(index), 0), clip_length - 1) for f in range(clip_length)]
        out.append(fi)
    return out


class OracleUNet:
    """UNetPseudo3DConditionModel.forward restated over a flat state dict (fp32, CPU)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], unet_config: dict, model_config: dict):
        self.w = {k: v.float() for k, v in state_dict.items()}
        self.cfg = dict(unet_config)
        self.mc = dict(model_config or {})
        self.heads = self.cfg["attention_head_dim"]
        self.groups = self.cfg["norm_num_groups"]
        self.eps = self.cfg["norm_eps"]
        self.ch = list(self.cfg["block_out_channels"])
        self.hook: Optional[Callable] = None

    def to(self, device):
        """Move the weights (tests only: the same restatement under CUDA fp16 autocast measures the reference's own fp16-vs-fp32 deviation)."""
        self.w = {k: v.to(device) for k, v in self.w.items()}
        return self

    # ---- primitives -----------------------------------------------------------------------------------------
    def _temporal(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """resnet.py:72-78 + lora.py:46-54 (LoRA: x + up(down(x)); else a Conv1d with bias)."""
        w = self.w
        b, c, f, h, wd = x.shape
        if name + ".conv_temporal.down.weight" in w:
            x1 = x.permute(0, 3, 4, 1, 2).reshape(b * h * wd, c, f)
            y = F.conv1d(x1, w[name + ".conv_temporal.down.weight"], padding=1)
            y = F.conv1d(y, w[name + ".conv_temporal.up.weight"], padding=1) + x1
        elif name + ".conv_temporal.weight" in w:
            x1 = x.permute(0, 3, 4, 1, 2).reshape(b * h * wd, c, f)
            k = w[name + ".conv_temporal.weight"].shape[-1]
            y = F.conv1d(x1, w[name + ".conv_temporal.weight"], w[name + ".conv_temporal.bias"], padding=k // 2)
        else:
            return x
        return y.reshape(b, h, wd, c, f).permute(0, 3, 4, 1, 2)

    def conv(self, name: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
        """PseudoConv3d.forward (resnet.py:57-80): per-frame 2-D conv, then the temporal conv for k>1."""
        wt = self.w[name + ".weight"]
        k = wt.shape[-1]
        b, c, f, h, wd = x.shape
        y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, wd), wt, self.w[name + ".bias"], stride=stride,
                     padding=k // 2)
        y = y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)
        return self._temporal(name, y) if k > 1 else y

    def linear(self, name: str, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.w[name + ".weight"], self.w.get(name + ".bias"))

    def gn5(self, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
        """GroupNorm on the 5-D tensor: statistics span (C/G, F, H, W) (resnet.py:338,369; unet_3d_condition.py:439)."""
        return F.group_norm(x, self.groups, self.w[name + ".weight"], self.w[name + ".bias"], eps)

    def ln(self, name: str, x: torch.Tensor) -> torch.Tensor:
        return F.layer_norm(x, (x.shape[-1],), self.w[name + ".weight"], self.w[name + ".bias"], 1e-5)

    # ---- blocks ---------------------------------------------------------------------------------------------
    def resnet(self, p: str, x: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
        """ResnetBlockPseudo3D.forward (resnet.py:335-394), time_embedding_norm='default'."""
        h = F.silu(self.gn5(p + ".norm1", x, self.eps))
        h = self.conv(p + ".conv1", h)
        t = self.linear(p + ".time_emb_proj", F.silu(temb))  # [B, Cout]
        h = h + t[:, :, None, None, None]
        h = F.silu(self.gn5(p + ".norm2", h, self.eps))
        h = self.conv(p + ".conv2", h)
        if p + ".conv_shortcut.weight" in self.w:
            x = self.conv(p + ".conv_shortcut", x)
        return x + h

    def _heads_to_batch(self, t: torch.Tensor) -> torch.Tensor:
        b, s, c = t.shape
        return t.reshape(b, s, self.heads, c // self.heads).permute(0, 2, 1, 3).reshape(b * self.heads, s, c // self.heads)

    def _attention(self, q, k, v, scale, is_cross: bool, place: str) -> torch.Tensor:
        """attention_register.py:23-59: softmax(scale*QK^T) -> controller([BF,heads,s,t]) -> PV."""
        scores = torch.baddbmm(q.new_empty(q.shape[0], q.shape[1], k.shape[1]), q, k.transpose(-1, -2), beta=0, alpha=scale)
        probs = scores.softmax(dim=-1)
        bh, s, t = probs.shape
        p4 = probs.reshape(bh // self.heads, self.heads, s, t)
        if self.hook is not None:
            p4 = self.hook(p4, is_cross, place)
        o = torch.bmm(p4.reshape(bh, s, t), v)
        o = o.reshape(bh // self.heads, self.heads, s, -1).permute(0, 2, 1, 3).reshape(bh // self.heads, s, -1)
        return o

    def st_attn(self, p: str, x: torch.Tensor, clip_length: int, index_list, place: str) -> torch.Tensor:
        """spatial_temporal_forward (attention_register.py:131-218)."""
        c = x.shape[-1]
        scale = (c // self.heads) ** -0.5
        q = self._heads_to_batch(self.linear(p + ".to_q", x))
        k = self.linear(p + ".to_k", x)
        v = self.linear(p + ".to_v", x)
        bf, s, _ = k.shape
        b = bf // clip_length
        k = k.reshape(b, clip_length, s, c)
        v = v.reshape(b, clip_length, s, c)
        if len(index_list) > 0:
            fis = sc_frame_indices(index_list, clip_length)
            k = torch.cat([k[:, fi] for fi in fis], dim=2)
            v = torch.cat([v[:, fi] for fi in fis], dim=2)
        k = self._heads_to_batch(k.reshape(bf, -1, c))
        v = self._heads_to_batch(v.reshape(bf, -1, c))
        o = self._attention(q, k, v, scale, False, place)
        return self.linear(p + ".to_out.0", o)

    def cross_attn(self, p: str, x: torch.Tensor, text: torch.Tensor, place: str) -> torch.Tensor:
        """forward (attention_register.py:71-128) with encoder_hidden_states."""
        c = x.shape[-1]
        scale = (c // self.heads) ** -0.5
        q = self._heads_to_batch(self.linear(p + ".to_q", x))
        k = self._heads_to_batch(self.linear(p + ".to_k", text))
        v = self._heads_to_batch(self.linear(p + ".to_v", text))
        o = self._attention(q, k, v, scale, True, place)
        return self.linear(p + ".to_out.0", o)

    def temporal_attn(self, p: str, x: torch.Tensor, clip_length: int) -> torch.Tensor:
        """apply_temporal_attention (models/attention.py:327-337): un-hooked CrossAttention over the frame axis."""
        bf, d, c = x.shape
        b = bf // clip_length
        h = x.reshape(b, clip_length, d, c).permute(0, 2, 1, 3).reshape(b * d, clip_length, c)
        n = self.ln(p + ".norm_temporal", h)
        scale = (c // self.heads) ** -0.5
        q = self._heads_to_batch(self.linear(p + ".attn_temporal.to_q", n))
        k = self._heads_to_batch(self.linear(p + ".attn_temporal.to_k", n))
        v = self._heads_to_batch(self.linear(p + ".attn_temporal.to_v", n))
        probs = torch.baddbmm(q.new_empty(q.shape[0], q.shape[1], k.shape[1]), q, k.transpose(-1, -2), beta=0,
                              alpha=scale).softmax(dim=-1)
        o = torch.bmm(probs, v)
        o = o.reshape(b * d, self.heads, clip_length, -1).permute(0, 2, 1, 3).reshape(b * d, clip_length, c)
        h = self.linear(p + ".attn_temporal.to_out.0", o) + h
        return h.reshape(b, d, clip_length, c).permute(0, 2, 1, 3).reshape(bf, d, c)

    def transformer(self, p: str, x: torch.Tensor, text: torch.Tensor, place: str) -> torch.Tensor:
        """SpatioTemporalTransformerModel.forward + Block.forward (models/attention.py:95-144,271-325)."""
        b, c, f, h, wd = x.shape
        xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, wd)
        text_f = text.repeat_interleave(f, 0)
        res = xf
        hs = F.group_norm(xf, self.groups, self.w[p + ".norm.weight"], self.w[p + ".norm.bias"], 1e-6)
        hs = F.conv2d(hs, self.w[p + ".proj_in.weight"], self.w[p + ".proj_in.bias"])
        hs = hs.permute(0, 2, 3, 1).reshape(b * f, h * wd, c)
        bp = p + ".transformer_blocks.0"
        # SparseCausalAttention_index selection (models/attention.py:169-173,292-293)
        if "SparseCausalAttention_index" in self.mc:
            index_list = list(self.mc["SparseCausalAttention_index"])
        else:
            index_list = [-1, "first"]
        if "least_sc_channel" in self.mc and c < self.mc["least_sc_channel"]:
            index_list = []
        hs = hs + self.st_attn(bp + ".attn1", self.ln(bp + ".norm1", hs), f, index_list, place)
        hs = self.cross_attn(bp + ".attn2", self.ln(bp + ".norm2", hs), text_f, place) + hs
        n3 = self.ln(bp + ".norm3", hs)
        proj = self.linear(bp + ".ff.net.0.proj", n3)
        a, gate = proj.chunk(2, dim=-1)
        hs = self.linear(bp + ".ff.net.2", a * F.gelu(gate)) + hs
        hs = self.temporal_attn(bp, hs, f)
        hs = hs.reshape(b * f, h, wd, c).permute(0, 3, 1, 2)
        hs = F.conv2d(hs, self.w[p + ".proj_out.weight"], self.w[p + ".proj_out.bias"])
        out = hs + res
        return out.reshape(b, f, c, h, wd).permute(0, 2, 1, 3, 4)

    def time_embedding(self, t, batch: int) -> torch.Tensor:
        """Timesteps(320, flip_sin_to_cos=True, freq_shift=0) + TimestepEmbedding (unet_3d_condition.py:356-362)."""
        c0 = self.ch[0]
        half = c0 // 2
        dev = self.w["time_embedding.linear_1.weight"].device
        ts = torch.full((batch,), float(t), dtype=torch.float32, device=dev)
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=dev) / (half - self.cfg.get("freq_shift", 0))
        emb = ts[:, None] * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.cfg.get("flip_sin_to_cos", True):
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        e = self.linear("time_embedding.linear_1", emb)
        return self.linear("time_embedding.linear_2", F.silu(e))

    # ---- forward ----------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, t, text: torch.Tensor, hook: Optional[Callable] = None) -> torch.Tensor:
        self.hook = hook
        x = x.float()
        text = text.float()
        temb = self.time_embedding(t, x.shape[0])
        h = self.conv("conv_in", x)
        skips = [h]
        nblk = len(self.ch)
        lpb = self.cfg["layers_per_block"]
        for i, btype in enumerate(self.cfg["down_block_types"]):
            p = f"down_blocks.{i}"
            for j in range(lpb):
                h = self.resnet(f"{p}.resnets.{j}", h, temb)
                if btype.startswith("CrossAttn"):
                    h = self.transformer(f"{p}.attentions.{j}", h, text, "down")
                skips.append(h)
            if i != nblk - 1:
                h = self.conv(f"{p}.downsamplers.0.conv", h, stride=2)
                skips.append(h)
        h = self.resnet("mid_block.resnets.0", h, temb)
        h = self.transformer("mid_block.attentions.0", h, text, "mid")
        h = self.resnet("mid_block.resnets.1", h, temb)
        for i, btype in enumerate(self.cfg["up_block_types"]):
            p = f"up_blocks.{i}"
            for j in range(lpb + 1):
                h = torch.cat([h, skips.pop()], dim=1)
                h = self.resnet(f"{p}.resnets.{j}", h, temb)
                if btype.startswith("CrossAttn"):
                    h = self.transformer(f"{p}.attentions.{j}", h, text, "up")
            if i != nblk - 1:
                b, c, f, hh, ww = h.shape
                u = F.interpolate(h.permute(0, 2, 1, 3, 4).reshape(b * f, c, hh, ww), scale_factor=2.0, mode="nearest")
                h = u.reshape(b, f, c, 2 * hh, 2 * ww).permute(0, 2, 1, 3, 4)
                h = self.conv(f"{p}.upsamplers.0.conv", h)
        h = F.silu(self.gn5("conv_norm_out", h, self.eps))
        out = self.conv("conv_out", h)
        self.hook = None
        return out


# --------------------------------------------------------------------------------------------------------------
# Host-side edit tables (ptp_utils.py:144-199, seq_aligner.py:61-195, attention_util.py:307-316)
# --------------------------------------------------------------------------------------------------------------


def word_token_indices(text: str, word_place, tokenizer) -> np.ndarray:
    """ptp_utils.get_word_inds: token positions (1-based, BOS at 0) of the word(s) selected by string or index."""
    words = text.split(" ")
    if isinstance(word_place, str):
        places = [i for i, w in enumerate(words) if w == word_place]
    elif isinstance(word_place, int):
        places = [word_place]
    else:
        places = list(word_place)
    out = []
    if places:
        pieces = [tokenizer.decode([tid]).strip("#") for tid in tokenizer.encode(text)][1:-1]
        cur_len, ptr = 0, 0
        for i, piece in enumerate(pieces):
            cur_len += len(piece)
            if ptr in places:
                out.append(i + 1)
            if cur_len >= len(words[ptr]):
                ptr += 1
                cur_len = 0
    return np.array(out)


def cross_replace_alpha_table(prompts, num_steps: int, cross_replace_steps, tokenizer, max_words: int = 77) -> torch.Tensor:
    """ptp_utils.get_time_words_attention_alpha -> [num_steps+1, 77] (single target prompt)."""
    crs = dict(cross_replace_steps) if isinstance(cross_replace_steps, dict) else {"default_": cross_replace_steps}
    if "default_" not in crs:
        crs["default_"] = (0.0, 1.0)
    n = num_steps + 1
    table = torch.zeros(n, len(prompts) - 1, max_words)

    def apply(bounds, prompt_ind, word_inds=None):
        if isinstance(bounds, float):
            bounds = (0, bounds)
        start, end = int(bounds[0] * n), int(bounds[1] * n)
        cols = torch.arange(max_words) if word_inds is None else torch.as_tensor(word_inds)
        table[:start, prompt_ind, cols] = 0
        table[start:end, prompt_ind, cols] = 1
        table[end:, prompt_ind, cols] = 0

    for i in range(len(prompts) - 1):
        apply(crs["default_"], i)
    for key, item in crs.items():
        if key == "default_":
            continue
        for i in range(1, len(prompts)):
            ind = word_token_indices(prompts[i], key, tokenizer)
            if len(ind) > 0:
                apply(item, i - 1, ind)
    return table[:, 0, :]


def _nw_align(x, y):
    """seq_aligner.global_align + get_aligned_sequences with ScoreParams(gap=0, match=1, mismatch=-1)."""
    nx, ny = len(x), len(y)
    score = np.zeros((nx + 1, ny + 1), dtype=np.int32)
    trace = np.zeros((nx + 1, ny + 1), dtype=np.int32)
    trace[0, 1:] = 1
    trace[1:, 0] = 2
    trace[0, 0] = 4
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            left = score[i, j - 1]
            up = score[i - 1, j]
            diag = score[i - 1, j - 1] + (1 if x[i - 1] == y[j - 1] else -1)
            best = max(left, up, diag)
            score[i, j] = best
            trace[i, j] = 1 if best == left else (2 if best == up else 3)
    pairs = []
    i, j = nx, ny
    while i > 0 or j > 0:
        tb = trace[i, j]
        if tb == 3:
            i -= 1
            j -= 1
            pairs.append((j, i))
        elif tb == 1:
            j -= 1
            pairs.append((j, -1))
        elif tb == 2:
            i -= 1
        else:
            break
    pairs.reverse()
    return pairs


def refinement_tables(prompts, tokenizer, max_len: int = 77):
    """seq_aligner.get_refinement_mapper -> (mapper[77] int64, alphas[77] float)."""
    x = tokenizer.encode(prompts[0])
    y = tokenizer.encode(prompts[1])
    pairs = _nw_align(x, y)
    base = torch.tensor(pairs, dtype=torch.int64)
    alphas = torch.ones(max_len)
    alphas[: base.shape[0]] = base[:, 1].ne(-1).float()
    mapper = torch.zeros(max_len, dtype=torch.int64)
    mapper[: base.shape[0]] = base[:, 1]
    mapper[base.shape[0]:] = len(y) + torch.arange(max_len - len(y))
    return mapper, alphas


def replacement_matrix(prompts, tokenizer, max_len: int = 77) -> torch.Tensor:
    """seq_aligner.get_replacement_mapper_ -> M[77,77] (source token w -> target token n)."""
    wx, wy = prompts[0].split(" "), prompts[1].split(" ")
    if len(wx) != len(wy):
        raise ValueError("attention replacement edit can only be applied on prompts with the same length")
    diff = [i for i in range(len(wy)) if wy[i] != wx[i]]
    src = [word_token_indices(prompts[0], i, tokenizer) for i in diff]
    tgt = [word_token_indices(prompts[1], i, tokenizer) for i in diff]
    m = np.zeros((max_len, max_len))
    i = j = 0
    cur = 0
    while i < max_len and j < max_len:
        if cur < len(src) and src[cur][0] == i:
            s_, t_ = src[cur], tgt[cur]
            if len(s_) == len(t_):
                m[s_, t_] = 1
            else:
                for it in t_:
                    m[s_, it] = 1 / len(t_)
            cur += 1
            i += len(s_)
            j += len(t_)
        elif cur < len(src):
            m[i, j] = 1
            i += 1
            j += 1
        else:
            m[j, j] = 1
            i += 1
            j += 1
    return torch.from_numpy(m).float()


def equalizer_row(text: str, words, values, tokenizer) -> torch.Tensor:
    """attention_util.get_equalizer -> [77]."""
    if isinstance(words, (int, str)):
        words = (words,)
    eq = torch.ones(77)
    for word, val in zip(words, values):
        eq[word_token_indices(text, word, tokenizer)] = val
    return eq


def blend_word_onehot(prompt: str, words, tokenizer) -> torch.Tensor:
    """SpatialBlender.__init__ (spatial_blend.py:138-147): one-hot [77] over the blend words of one prompt."""
    a = torch.zeros(77)
    if isinstance(words, str):
        words = [words]
    for word in words:
        ind = word_token_indices(prompt, word, tokenizer)
        a[ind] = 1
    return a


class EditPlan:
    """Everything `make_controller` (attention_util.py:320-387) precomputes, as plain tensors."""

    def __init__(self, tokenizer, source_prompt: str, target_prompt: str, num_steps: int, cross_replace_steps,
                 self_replace_steps, is_replace_controller: bool = True, eq_params=None, blend_words=None,
                 blend_self_attention: bool = False, blend_latents: bool = False, blend_th=(0.3, 0.3)):
        prompts = [source_prompt, target_prompt]
        self.N = num_steps
        # p2p_ddim_spatial_temporal.py:176-184: Replace only when word counts match
        equal_length = len(source_prompt.split(" ")) == len(target_prompt.split(" "))
        self.mode = "replace" if (is_replace_controller and equal_length) else "refine"
        self.alpha = cross_replace_alpha_table(prompts, num_steps, cross_replace_steps, tokenizer)  # [N+1,77]
        srs = self_replace_steps
        if isinstance(srs, float):
            srs = (0, srs)
        self.self_window = (int(num_steps * srs[0]), int(num_steps * srs[1]))
        if self.mode == "replace":
            self.M = replacement_matrix(prompts, tokenizer)
        else:
            self.mapper, self.a = refinement_tables(prompts, tokenizer)
        self.eq = None
        if eq_params is not None:
            self.eq = equalizer_row(target_prompt, eq_params["words"], eq_params["values"], tokenizer)
        self.blend_src = self.blend_tgt = None
        self.blend_self = self.blend_lat = False
        if blend_words is not None and blend_words != "None":
            self.blend_src = blend_word_onehot(source_prompt, blend_words[0], tokenizer)
            self.blend_tgt = blend_word_onehot(target_prompt, blend_words[1], tokenizer)
            self.blend_self = bool(blend_self_attention)
            self.blend_lat = bool(blend_latents)
        self.th = tuple(blend_th)
        self.lat_window = (int(0.2 * num_steps), int(0.8 * num_steps))


# --------------------------------------------------------------------------------------------------------------
# Controllers (attention_store.py, attention_util.py, spatial_blend.py) as small state machines
# --------------------------------------------------------------------------------------------------------------

_KEYS = ("down_cross", "mid_cross", "up_cross", "down_self", "mid_self", "up_self")


def _empty_store():
    return {k: [] for k in _KEYS}


class OracleStore:
    """AttentionStore in LOW_RESOURCE mode (inversion): keeps every map with <=1024 queries per step."""

    def __init__(self, save_self_attention: bool = True):
        self.save_self = save_self_attention
        self.step_store = _empty_store()
        self.attention_store: Dict[str, List[torch.Tensor]] = {}
        self.all_step: List[Dict[str, List[torch.Tensor]]] = []
        self.latents_store: List[torch.Tensor] = []
        self.cur_step = 0

    def store(self, attn: torch.Tensor, is_cross: bool, place: str):
        if attn.shape[-2] <= 32 ** 2 and (is_cross or self.save_self):
            self.step_store[f"{place}_{'cross' if is_cross else 'self'}"].append(attn.clone())

    def hook(self, attn, is_cross, place):  # AttentionControl.__call__ with LOW_RESOURCE=True
        self.store(attn, is_cross, place)
        return attn

    def between_steps(self):
        if not self.attention_store:
            self.attention_store = {k: [t.clone() for t in v] for k, v in self.step_store.items()}
        else:
            for k in self.attention_store:
                for i in range(len(self.attention_store[k])):
                    self.attention_store[k][i] = self.attention_store[k][i] + self.step_store[k][i]
        self.all_step.append(self.step_store)
        self.step_store = _empty_store()

    def step_callback(self, x_t):
        self.cur_step += 1
        self.between_steps()
        self.latents_store.append(x_t.clone())
        return x_t


def blend_mask(maps5: List[torch.Tensor], word_alpha: torch.Tensor, th: float, h: int, w: int) -> torch.Tensor:
    """SpatialBlender.get_mask with use_pool=True (spatial_blend.py:24-39) for ONE prompt.

    maps5: the stored cross maps down_cross[2:4] + up_cross[:3], each [F, heads, r*r, 77]. Returns bool [F, h, w]."""
    stack = []
    for m in maps5:
        f, heads, rr, _ = m.shape
        r = int(math.isqrt(rr))
        stack.append((m.float() * word_alpha[None, None, None, :]).sum(-1).reshape(f, heads, r, r).permute(1, 0, 2, 3))
    maps = torch.cat(stack, dim=0).mean(0)  # [F, r, r]   (mean over layers*heads)
    maps = F.max_pool2d(maps[None], (3, 3), (1, 1), padding=(1, 1))
    mask = F.interpolate(maps, size=(h, w))[0]
    mask = mask / mask.amax(dim=(-2, -1), keepdim=True)
    return mask.gt(th)


class OracleEdit(OracleStore):
    """AttentionControlEdit (+Replace / Refine / Reweight) with an external inversion store (use_inversion_attention)."""

    def __init__(self, plan: EditPlan, inv: OracleStore, save_self_attention: bool = False):
        super().__init__(save_self_attention)
        self.plan = plan
        self.inv = inv
        self.pos = {k: 0 for k in _KEYS}
        self.lat_counter = 0
        self.mask_list: List[torch.Tensor] = []

    def _blend_maps(self, store_dict):
        return store_dict["down_cross"][2:4] + store_dict["up_cross"][:3]

    def hook(self, attn, is_cross, place):  # AttentionControl.__call__ with CFG: only the cond half is edited
        h = attn.shape[0]
        attn = attn.clone()
        attn[h // 2:] = self.forward(attn[h // 2:], is_cross, place)
        return attn

    def forward(self, attn, is_cross, place):
        pl = self.plan
        self.store(attn, is_cross, place)  # pre-edit target maps accumulate (App. B quirk 4)
        if attn.shape[-2] > 32 ** 2:
            return attn
        key = f"{place}_{'cross' if is_cross else 'self'}"
        pos = self.pos[key]
        self.pos[key] += 1
        step_in_store = len(self.inv.all_step) - self.cur_step - 1
        base = self.inv.all_step[step_in_store][key][pos]
        k = self.cur_step
        if is_cross:
            if pl.mode == "refine":
                r = base[..., pl.mapper] * pl.a + attn * (1 - pl.a)
            else:
                r = torch.einsum("thpw,wn->thpn", base, pl.M)
            if pl.eq is not None:
                r = r * pl.eq
            al = pl.alpha[k]
            return r * al + (1 - al) * attn
        if pl.self_window[0] <= k < pl.self_window[1]:
            if pl.blend_self:
                rr = attn.shape[-2]
                r = int(math.isqrt(rr))
                m = blend_mask(self._blend_maps(self.inv.all_step[step_in_store]), pl.blend_src, pl.th[0], r, r)
                m = m.float().reshape(m.shape[0], 1, rr, 1)
                return m * attn + (1 - m) * base
            return base
        return attn

    def between_steps(self):
        super().between_steps()
        self.pos = {k: 0 for k in _KEYS}

    def step_callback(self, x_t):
        """AttentionControlEdit.step_callback (attention_util.py:47-78): bookkeeping then optional latent blend."""
        self.cur_step += 1
        self.between_steps()
        self.latents_store.append(x_t.clone())
        pl = self.plan
        if not pl.blend_lat:
            return x_t
        step_in_store = len(self.inv.latents_store) - self.cur_step
        x_inv = self.inv.latents_store[step_in_store]
        src = self._blend_maps(self.inv.all_step[step_in_store])
        tgt = self._blend_maps(self.attention_store)  # running sums of the pre-edit target maps
        h, w = x_t.shape[-2:]
        m_src = blend_mask(src, pl.blend_src, pl.th[0], h, w)
        m_tgt = blend_mask(tgt, pl.blend_tgt, pl.th[0], h, w)
        m = (m_src | m_tgt).float()  # 'both': mask[:1] + mask, row 1 (spatial_blend.py:40-42)
        self.mask_list.append(m_src.float())  # reference appends mask[0] (= source mask)
        self.lat_counter += 1
        if pl.lat_window[0] < self.lat_counter < pl.lat_window[1]:
            return x_inv + m[None, None] * (x_t - x_inv)
        return x_t


# --------------------------------------------------------------------------------------------------------------
# DDIM loops (p2p_ddim_spatial_temporal.py:131-161, 389-413)
# --------------------------------------------------------------------------------------------------------------


def invert(unet: OracleUNet, x0: torch.Tensor, cond_text: torch.Tensor, num_steps: int, store: Optional[OracleStore]):
    """ddim_clean2noisy_loop + next_clean2noise_step. Returns the N+1 latents (x0 first)."""
    ac = alphas_cumprod()
    ts = ddim_timesteps(num_steps)
    x = x0.clone().float()
    out = [x0.clone().float()]
    for i in range(num_steps):
        t = ts[num_steps - 1 - i]
        eps = unet.forward(x, t, cond_text, store.hook if store is not None else None)
        t_prev = min(t - 1000 // num_steps, 999)
        a_prev = ac[t_prev] if t_prev >= 0 else ac[0]
        a_next = ac[t]
        x0_pred = (x - (1 - a_prev) ** 0.5 * eps) / a_prev ** 0.5
        x = a_next ** 0.5 * x0_pred + (1 - a_next) ** 0.5 * eps
        if store is not None:
            store.step_callback(x)
        out.append(x.clone())
    return out


def edit(unet: OracleUNet, x_T: torch.Tensor, text_uc: torch.Tensor, num_steps: int, ctrl: Optional[OracleEdit],
         guidance_scale: float = 7.5):
    """sd_ddim_pipeline step loop with CFG and DDIMScheduler.step(eta=0). text_uc = cat(uncond, cond) [2,77,D]."""
    ac = alphas_cumprod()
    ts = ddim_timesteps(num_steps)
    x = x_T.clone().float()
    trace = []
    for t in ts:
        x2 = torch.cat([x, x], dim=0)
        eps2 = unet.forward(x2, t, text_uc, ctrl.hook if ctrl is not None else None)
        e_u, e_c = eps2.chunk(2)
        eps = e_u + guidance_scale * (e_c - e_u)
        t_prev = t - 1000 // num_steps
        a_t = ac[t]
        a_prev = ac[t_prev] if t_prev >= 0 else ac[0]
        x0_pred = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        x = a_prev ** 0.5 * x0_pred + (1 - a_prev) ** 0.5 * eps
        if ctrl is not None:
            x = ctrl.step_callback(x)
        trace.append(x.clone())
    return trace


def encode_prompts(tokenizer, text_encoder, prompt: str) -> torch.Tensor:
    """stable_diffusion.py:180-295 with CFG: cat(uncond("") , cond(prompt)) -> [2,77,D]."""
    ids = tokenizer(prompt, padding="max_length", max_length=tokenizer.model_max_length, truncation=True).input_ids
    un = tokenizer([""], padding="max_length", max_length=ids.shape[-1], truncation=True).input_ids
    with torch.no_grad():
        return torch.cat([text_encoder(un)[0], text_encoder(ids)[0]]).float()
