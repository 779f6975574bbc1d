"""Thin torch-tensor wrappers over the C ABI (include/fatezero_b200.h).  PyTorch is plumbing here: device memory,
streams, and nothing else — every compute call below lands in libfatezero_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AttnArgs, Epilogue

f16 = torch.float16


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"fatezero_b200.ops.{name}: tensor must live on a CUDA device (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"fatezero_b200.ops.{name}: expected {dtype}, got {t.dtype}")


def geglu_block_n(gemm_cols: int) -> int:
    for bn in (256, 160, 128, 64):
        if gemm_cols % bn == 0:
            return bn
    return 32


def pack_geglu(weight: torch.Tensor, bias: Optional[torch.Tensor]):
    """[2*Nout, K] (x rows then gate rows, diffusers GEGLU.proj) -> tile-interleaved so each BLOCK_N tile holds x|gate."""
    two_n, k = weight.shape
    nout = two_n // 2
    bn = geglu_block_n(two_n)
    half = bn // 2
    assert nout % half == 0
    wx, wg = weight[:nout].reshape(nout // half, half, k), weight[nout:].reshape(nout // half, half, k)
    w = torch.cat([wx, wg], dim=1).reshape(two_n, k).contiguous()
    b = None
    if bias is not None:
        bx, bg = bias[:nout].reshape(nout // half, half), bias[nout:].reshape(nout // half, half)
        b = torch.cat([bx, bg], dim=1).reshape(two_n).contiguous()
    return w, b, bn


def _epilogue(bias=None, group_bias=None, rows_per_group=0, residual=None, geglu=False, residual2=None, vt_out=None, vt_col_start=0, vt_S=0, vt_d=0,
              vt_heads=0, vt_ld=0) -> Epilogue:
    e = Epilogue()
    e.bias = _p(bias)
    e.group_bias = _p(group_bias)
    e.rows_per_group = int(rows_per_group)
    e.residual = _p(residual)
    e.ldr = residual.stride(-2) if residual is not None else 0
    e.residual2 = _p(residual2)
    e.ldr2 = residual2.stride(-2) if residual2 is not None else 0
    e.mode = _lib.EPI_GEGLU if geglu else _lib.EPI_ROWMAJOR
    e.out_vt = _p(vt_out)
    e.vt_col_start = int(vt_col_start)
    e.vt_S, e.vt_d, e.vt_heads, e.vt_ld = int(vt_S), int(vt_d), int(vt_heads), int(vt_ld)
    return e


def gemm(a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, out=None, group_bias=None, rows_per_group=0, geglu=False,
         n_out: Optional[int] = None, vt: Optional[dict] = None, force_bn: int = 0) -> torch.Tensor:
    """out[M, N] = a[M, K] @ w[N, K]^T (+bias +group_bias +residual) ; geglu: w packed by pack_geglu, N_out = N/2."""
    _chk(a, f16, "gemm"); _chk(w, f16, "gemm")
    M, K = a.shape
    N = w.shape[0]
    assert a.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K
    cols_out = N // 2 if geglu else N
    if vt is not None:
        cols_out = vt["col_start"]
    if n_out is not None:
        cols_out = n_out
    if out is None:
        out = torch.empty((M, max(cols_out, 8)), dtype=f16, device=a.device)
    e = _epilogue(bias, group_bias, rows_per_group, residual, geglu, **({} if vt is None else dict(
        vt_out=vt["out"], vt_col_start=vt["col_start"], vt_S=vt["S"], vt_d=vt["d"], vt_heads=vt["heads"], vt_ld=vt.get("ld", 0))))
    _lib.call("fz_gemm_f16", _p(a), a.stride(0), _p(w), w.stride(0), M, N, K, C.byref(e), _p(out), out.stride(0), force_bn, _stream())
    return out


def conv3x3(x: torch.Tensor, w9: torch.Tensor, bias=None, stride: int = 1, residual=None, group_bias=None, rows_per_group=0,
            force_bn: int = 0, asym_pad: bool = False) -> torch.Tensor:
    """x [NB,H,W,Cin] fp16 NHWC, w9 [9,Cout,Cin] -> [NB,H/stride,W/stride,Cout].  asym_pad (stride 2): right/bottom-only padding."""
    _chk(x, f16, "conv3x3"); _chk(w9, f16, "conv3x3")
    NB, H, W, Cin = x.shape
    Cout = w9.shape[1]
    assert x.is_contiguous() and w9.is_contiguous() and w9.shape == (9, Cout, Cin)
    out = torch.empty((NB, H // stride, W // stride, Cout), dtype=f16, device=x.device)
    e = _epilogue(bias, group_bias, rows_per_group, residual)
    if residual is not None:
        e.ldr = residual.shape[-1]
    if asym_pad:
        assert stride == 2
        _lib.call("fz_conv3x3_down_asym_nhwc_f16", _p(x), Cin, NB, H, W, Cin, _p(w9), Cout, C.byref(e), _p(out), Cout, force_bn, _stream())
    else:
        _lib.call("fz_conv3x3_nhwc_f16", _p(x), Cin, NB, H, W, Cin, _p(w9), Cout, stride, C.byref(e), _p(out), Cout, force_bn, _stream())
    return out


def tconv3(x: torch.Tensor, w3: torch.Tensor, bias=None, residual=None, group_bias=None, rows_per_group=0, force_bn: int = 0,
           residual2=None, halo: bool = False):
    """x [B,F,HW,Cin], w3 [3,Cout,Cin]: Conv1d(k=3,pad=1) over F -> [B,F,HW,Cout] (+bias +residual +group_bias).
    halo (frame-sharded execution): x is [B,F+2,HW,Cin] with the neighbour ranks' boundary frames in frames 0 and F+1."""
    _chk(x, f16, "tconv3"); _chk(w3, f16, "tconv3")
    B, F, HW, Cin = x.shape
    if halo:
        F -= 2
    Cout = w3.shape[1]
    assert x.is_contiguous() and w3.is_contiguous()
    out = torch.empty((B, F, HW, Cout), dtype=f16, device=x.device)
    e = _epilogue(bias, group_bias, rows_per_group, residual, residual2=residual2)
    if residual is not None:
        e.ldr = residual.shape[-1]
    if residual2 is not None:
        e.ldr2 = residual2.shape[-1]
    _lib.call("fz_tconv3_halo_f16" if halo else "fz_tconv3_f16", _p(x), Cin, B, F, HW, Cin, _p(w3), Cout, C.byref(e), _p(out), Cout, force_bn,
              _stream())
    return out


_gn_ws = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    ws = _gn_ws.get(device)
    if ws is None or ws.numel() * 8 < nbytes:
        ws = torch.zeros(max(nbytes // 8 + 1, 4096), dtype=torch.float64, device=device)  # arrival counters start at zero
        _gn_ws[device] = ws
    return ws


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, groups: int, frames_per_stat: int, silu: bool,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [NB, HW, C] fp16; statistics over (C/groups, HW, frames_per_stat consecutive NB rows)."""
    _chk(x, f16, "groupnorm")
    NB, HW, Cc = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    ws = _workspace(x.device, 1 << 20)  # per-(image, chunk, group) partial sums
    _lib.call("fz_groupnorm_nhwc_f16", _p(x), _p(out), NB, HW, Cc, groups, frames_per_stat, _p(gamma), _p(beta), float(eps), int(silu),
              _p(ws), _stream())
    return out


def groupnorm_stats(x: torch.Tensor, groups: int) -> torch.Tensor:
    """Per-image (sum, sumsq) of every group: fp32 view [NB, groups, 2] INTO the shared workspace (valid until the next groupnorm call)."""
    _chk(x, f16, "groupnorm_stats")
    NB, HW, Cc = x.shape
    assert x.is_contiguous()
    ws = _workspace(x.device, 1 << 20)
    _lib.call("fz_groupnorm_stats_f16", _p(x), NB, HW, Cc, groups, _p(ws), _stream())
    off = (768 * 1024) // 4
    return ws.view(torch.float32)[off: off + NB * groups * 2].view(NB, groups, 2)


def groupnorm_apply(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, groups: int, frames_per_stat: int, count_frames: int,
                    silu: bool, sums: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Second half of the frame-sharded GroupNorm: sums [NB, groups, 2] fp32 (the apply adds frames_per_stat consecutive images)."""
    _chk(x, f16, "groupnorm_apply")
    NB, HW, Cc = x.shape
    assert x.is_contiguous() and sums.is_contiguous() and sums.dtype == torch.float32 and sums.numel() == NB * groups * 2
    if out is None:
        out = torch.empty_like(x)
    _lib.call("fz_groupnorm_apply_f16", _p(x), _p(out), NB, HW, Cc, groups, frames_per_stat, count_frames, _p(gamma), _p(beta), float(eps),
              int(silu), _p(sums), _stream())
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _chk(x, f16, "layernorm")
    M, Cc = x.shape
    assert x.is_contiguous()
    out = torch.empty_like(x)
    _lib.call("fz_layernorm_f16", _p(x), _p(out), M, Cc, _p(gamma), _p(beta), float(eps), _stream())
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    _chk(x, f16, "upsample2x")
    NB, H, W, Cc = x.shape
    out = torch.empty((NB, 2 * H, 2 * W, Cc), dtype=f16, device=x.device)
    _lib.call("fz_upsample2x_nhwc_f16", _p(x), _p(out), NB, H, W, Cc, _stream())
    return out


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _chk(a, f16, "concat"); _chk(b, f16, "concat")
    rows = a.numel() // a.shape[-1]
    out = torch.empty((*a.shape[:-1], a.shape[-1] + b.shape[-1]), dtype=f16, device=a.device)
    _lib.call("fz_concat_channels_f16", _p(a), a.shape[-1], _p(b), b.shape[-1], _p(out), rows, _stream())
    return out


def im2col_latents(x: torch.Tensor) -> torch.Tensor:
    """latents [B,Cl,F,H,W] fp32 -> [B*F*H*W, 64] fp16."""
    _chk(x, torch.float32, "im2col_latents")
    B, Cl, F, H, W = x.shape
    out = torch.empty((B * F * H * W, 64), dtype=f16, device=x.device)
    _lib.call("fz_im2col_latents_f16", _p(x.contiguous()), _p(out), B, Cl, F, H, W, _stream())
    return out


def out_temporal(y: torch.Tensor, B: int, Co: int, F: int, H: int, W: int, down=None, up=None, w_full=None, b_full=None) -> torch.Tensor:
    """y [B*F*H*W, ld] fp16 (Co valid) -> eps [B,Co,F,H,W] fp32 with the conv_out temporal conv applied."""
    eps = torch.empty((B, Co, F, H, W), dtype=torch.float32, device=y.device)
    rank = 0 if down is None else down.shape[0]
    _lib.call("fz_out_temporal_f32", _p(y), y.stride(0), _p(eps), B, Co, F, H * W, _p(down), _p(up), rank, _p(w_full), _p(b_full), _stream())
    return eps


def rowvec_linear(x: torch.Tensor, w16: torch.Tensor, bias: Optional[torch.Tensor], silu_in: bool) -> torch.Tensor:
    N, K = w16.shape
    y = torch.empty((N,), dtype=torch.float32, device=x.device)
    _lib.call("fz_rowvec_linear", _p(x), _p(w16), _p(bias), _p(y), N, K, int(silu_in), _stream())
    return y


def timestep_sinusoid(t: float, c0: int, flip: bool, freq_shift: float, device) -> torch.Tensor:
    out = torch.empty((c0,), dtype=torch.float32, device=device)
    _lib.call("fz_timestep_sinusoid", float(t), _p(out), c0, int(flip), float(freq_shift), _stream())
    return out


def temporal_attn(qkv: torch.Tensor, B: int, F: int, HW: int, heads: int, d: int, scale: float) -> torch.Tensor:
    _chk(qkv, f16, "temporal_attn")
    out = torch.empty((B * F * HW, heads * d), dtype=f16, device=qkv.device)
    _lib.call("fz_temporal_attn_f16", _p(qkv), _p(out), B, F, HW, heads, d, float(scale), _stream())
    return out


def ddim_invert_step(x: torch.Tensor, eps: torch.Tensor, a_prev: float, a_next: float):
    _lib.call("fz_ddim_invert_step", _p(x), _p(eps), x.numel(), float(a_prev), float(a_next), _stream())


def cfg_ddim_step(x: torch.Tensor, eps2: torch.Tensor, guidance: float, a_t: float, a_prev: float, x_inv=None, mask_a=None, mask_b=None,
                  apply_blend: bool = False):
    fhw = x.shape[-3] * x.shape[-2] * x.shape[-1]
    _lib.call("fz_cfg_ddim_step", _p(x), _p(eps2), x.numel(), float(guidance), float(a_t), float(a_prev), _p(x_inv), _p(mask_a), _p(mask_b),
              fhw, int(apply_blend), _stream())


def blend_mask(maps: Sequence[torch.Tensor], word_w: torch.Tensor, th: float, h: int, w: int) -> torch.Tensor:
    """maps: list of [F, heads, r*r, ld] (fp16 cache slabs or fp16 running sums) -> mask [F, h, w] float 0/1."""
    m0 = maps[0]
    Fm, heads, rr, ld = m0.shape
    r = int(round(rr ** 0.5))
    arr = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
    ww = [float(v) for v in word_w.tolist()]
    wv = (C.c_float * len(ww))(*ww)
    out = torch.empty((Fm, h, w), dtype=torch.float32, device=m0.device)
    _lib.call("fz_blend_mask", arr, len(maps), int(m0.dtype == torch.float32), Fm, heads, r, m0.stride(2), min(len(ww), 77), wv,
              float(th), h, w, _p(out), _stream())
    return out


def softmax_rows_(x: torch.Tensor, scale: float) -> torch.Tensor:
    """x [rows, n] fp16 (row stride multiple of 8) <- softmax(scale * x) row-wise, in place."""
    _chk(x, f16, "softmax_rows")
    rows, n = x.shape
    assert x.stride(1) == 1
    _lib.call("fz_softmax_rows_f16", _p(x), rows, n, x.stride(0), float(scale), _stream())
    return x


def embed_tokens(tok: torch.Tensor, pos: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """tok [V, C] fp32, pos [L, C] fp32, ids [B, L] int64 (CUDA) -> [B*L, C] fp16."""
    B, L = ids.shape
    out = torch.empty((B * L, tok.shape[1]), dtype=f16, device=tok.device)
    _lib.call("fz_embed_tokens_f16", _p(tok), _p(pos), _p(ids.contiguous()), _p(out), B * L, L, tok.shape[1], _stream())
    return out


def quick_gelu_(x: torch.Tensor) -> torch.Tensor:
    _chk(x, f16, "quick_gelu")
    assert x.is_contiguous()
    _lib.call("fz_quick_gelu_f16", _p(x), x.numel(), _stream())
    return x


def cross_heatmaps(maps: Sequence[torch.Tensor], ntok: int) -> torch.Tensor:
    """maps: cross-attention running sums [F, heads, r*r, ld] (fp16 or fp32) of ONE resolution -> uint8 [F, ntok, r, r] heat maps."""
    m0 = maps[0]
    Fm, heads, rr, _ = m0.shape
    r = int(round(rr ** 0.5))
    arr = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
    out = torch.empty((Fm, ntok, r, r), dtype=torch.uint8, device=m0.device)
    _lib.call("fz_cross_heatmaps", arr, len(maps), int(m0.dtype == torch.float32), Fm, heads, r, m0.stride(2), int(ntok), _p(out), _stream())
    return out


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, *, S_q: int, keys_per_slot: int, n_src: int, d: int,
              heads: int, F: int, BF: int, scale: float, src_index: Sequence[Sequence[int]], edit_bf_start: int = 0,
              row_mode: int = _lib.ATTN_NONE, store=None, base=None, cache_ld: int = 0, acc=None, xedit=None, mask=None, dbg=None,
              causal: bool = False):
    """q/k: strided 2-D views (rows, ld) whose column h*d starts head h; vt [n_src, heads, d, vt_ld]; out [BF*S_q, ldo]."""
    a = AttnArgs()
    a.q, a.ldq = _p(q), q.stride(0)
    a.k, a.ldk = _p(k), k.stride(0)
    a.vt, a.vt_ld = _p(vt), vt.stride(2)
    a.out, a.ldo = _p(out), out.stride(0)
    a.S_q, a.keys_per_slot, a.n_slots, a.n_src = S_q, keys_per_slot, len(src_index), n_src
    a.d, a.heads, a.F, a.BF = d, heads, F, BF
    a.scale = float(scale)
    flat = [int(v) for row in src_index for v in row]
    arr = (C.c_int * len(flat))(*flat)
    a.src_index = arr
    a.edit_bf_start, a.row_mode = edit_bf_start, row_mode
    a.store, a.base, a.cache_ld = _p(store), _p(base), cache_ld
    a.acc, a.acc_ld = _p(acc), (acc.stride(2) if acc is not None else 0)
    a.xedit, a.mask = _p(xedit), _p(mask)
    a.dbg = _p(dbg)
    a.causal = int(causal)
    _lib.call("fz_attention_f16", C.byref(a), _stream())
    return out
