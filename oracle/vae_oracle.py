"""TEST INFRASTRUCTURE — fp32 torch restatement of the SD-1.x VAE (`AutoencoderKL` of diffusers 0.11.1: `models/vae.py`, `models/unet_2d_blocks.py`
{DownEncoderBlock2D, UpDecoderBlock2D, UNetMidBlock2D}, `models/resnet.py` {ResnetBlock2D, Downsample2D, Upsample2D}, `models/attention.py`
AttentionBlock), which the reference calls at pipelines/p2p_ddim_spatial_temporal.py:88-96 (encode) and pipelines/stable_diffusion.py:297-319
(decode).

PARITY UNPINNED: diffusers is not vendored in /root/reference and no wheel is available offline, and the reference has no
test or golden vector for the VAE, so this file restates the published architecture from the state-dict layout and forward order of that
release; it has NOT been diffed against the real package.  It is the checker for fatezero_b200/vae.py, never part of the product.
State-dict names follow diffusers 0.11.1 (`mid_block.attentions.0.{group_norm,query,key,value,proj_attn}`)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD14_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4,
                       norm_num_groups=32)


def vae_param_spec(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """name -> shape of every tensor of the AutoencoderKL state dict."""
    ch = list(cfg["block_out_channels"])
    lpb, lat = cfg["layers_per_block"], cfg["latent_channels"]
    spec: Dict[str, Tuple[int, ...]] = {}

    def conv(n, ci, co, k=3):
        spec[n + ".weight"] = (co, ci, k, k)
        spec[n + ".bias"] = (co,)

    def norm(n, c):
        spec[n + ".weight"] = (c,)
        spec[n + ".bias"] = (c,)

    def lin(n, ci, co):
        spec[n + ".weight"] = (co, ci)
        spec[n + ".bias"] = (co,)

    def resnet(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", ci, co); norm(n + ".norm2", co); conv(n + ".conv2", co, co)
        if ci != co:
            conv(n + ".conv_shortcut", ci, co, 1)

    def mid(n, c):
        resnet(n + ".resnets.0", c, c)
        norm(n + ".attentions.0.group_norm", c)
        for q in ("query", "key", "value", "proj_attn"):
            lin(f"{n}.attentions.0.{q}", c, c)
        resnet(n + ".resnets.1", c, c)

    conv("encoder.conv_in", cfg["in_channels"], ch[0])
    co = ch[0]
    for i, c in enumerate(ch):
        ci, co = co, c
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if i != len(ch) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co)
    mid("encoder.mid_block", ch[-1])
    norm("encoder.conv_norm_out", ch[-1])
    conv("encoder.conv_out", ch[-1], 2 * lat)
    conv("quant_conv", 2 * lat, 2 * lat, 1)
    conv("post_quant_conv", lat, lat, 1)
    rev = list(reversed(ch))
    conv("decoder.conv_in", lat, rev[0])
    mid("decoder.mid_block", rev[0])
    co = rev[0]
    for i, c in enumerate(rev):
        ci, co = co, c
        for j in range(lpb + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if i != len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co)
    norm("decoder.conv_norm_out", rev[-1])
    conv("decoder.conv_out", rev[-1], cfg["out_channels"])
    return spec


class VaeOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: dict = SD14_VAE_CONFIG):
        self.w = {k: v.float() for k, v in state_dict.items()}
        self.cfg = dict(cfg)
        self.groups = cfg["norm_num_groups"]

    def to(self, device):
        self.w = {k: v.to(device) for k, v in self.w.items()}
        return self

    def _conv(self, n, x, stride=1, padding=1):
        return F.conv2d(x, self.w[n + ".weight"], self.w[n + ".bias"], stride=stride, padding=padding)

    def _gn(self, n, x):
        return F.group_norm(x, self.groups, self.w[n + ".weight"], self.w[n + ".bias"], 1e-6)

    def _resnet(self, n, x):
        h = self._conv(n + ".conv1", F.silu(self._gn(n + ".norm1", x)))
        h = self._conv(n + ".conv2", F.silu(self._gn(n + ".norm2", h)))
        if n + ".conv_shortcut.weight" in self.w:
            x = self._conv(n + ".conv_shortcut", x, padding=0)
        return x + h

    def _attn(self, n, x):
        b, c, hh, ww = x.shape
        h = self._gn(n + ".group_norm", x).reshape(b, c, hh * ww).transpose(1, 2)
        q = F.linear(h, self.w[n + ".query.weight"], self.w[n + ".query.bias"])
        k = F.linear(h, self.w[n + ".key.weight"], self.w[n + ".key.bias"])
        v = F.linear(h, self.w[n + ".value.weight"], self.w[n + ".value.bias"])
        scale = 1.0 / math.sqrt(math.sqrt(c))  # one head of width c: both operands are scaled (AttentionBlock.forward)
        p = torch.softmax((q * scale) @ (k * scale).transpose(1, 2), dim=-1)
        o = F.linear(p @ v, self.w[n + ".proj_attn.weight"], self.w[n + ".proj_attn.bias"])
        return o.transpose(1, 2).reshape(b, c, hh, ww) + x

    def _mid(self, n, x):
        x = self._resnet(n + ".resnets.0", x)
        x = self._attn(n + ".attentions.0", x)
        return self._resnet(n + ".resnets.1", x)

    @torch.no_grad()
    def encode_moments(self, img: torch.Tensor) -> torch.Tensor:
        """img [N, 3, H, W] in [-1, 1] -> moments [N, 2*latent, H/8, W/8] (mean | logvar), AutoencoderKL.encode before the sampling."""
        ch, lpb = self.cfg["block_out_channels"], self.cfg["layers_per_block"]
        x = self._conv("encoder.conv_in", img)
        for i in range(len(ch)):
            for j in range(lpb):
                x = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", x)
            if i != len(ch) - 1:
                x = self._conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)
        x = self._mid("encoder.mid_block", x)
        x = self._conv("encoder.conv_out", F.silu(self._gn("encoder.conv_norm_out", x)))
        return self._conv("quant_conv", x, padding=0)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        ch, lpb = self.cfg["block_out_channels"], self.cfg["layers_per_block"]
        x = self._conv("decoder.conv_in", self._conv("post_quant_conv", z, padding=0))
        x = self._mid("decoder.mid_block", x)
        for i in range(len(ch)):
            for j in range(lpb + 1):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x)
            if i != len(ch) - 1:
                x = self._conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return self._conv("decoder.conv_out", F.silu(self._gn("decoder.conv_norm_out", x)))
