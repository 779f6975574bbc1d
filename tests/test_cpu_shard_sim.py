"""The host logic of the frame-sharded forward on CPU, world sizes 2..8: the engine's own exchange methods
(`UNetEngine._halo_ext`, `_kv_exchange`, `_temporal_attn_sharded`, `_finish`) run unmodified, one Python thread per simulated rank, against a
fake symmetric arena whose `exchange()` performs the 2-D segment copies with memmove into the peers' host buffers and then meets the other
ranks at a barrier (the push + flag + wait of fz_p2p_push).  What is checked is the index arithmetic the CUDA kernels are fed with: which
bytes of which rank land where, and that the sharded results equal the un-sharded computation on the whole clip."""
import ctypes
import threading

import pytest
import torch

from fatezero_b200 import engine as eng_mod
from fatezero_b200 import ops
from fatezero_b200.engine import UNetEngine, sc_frame_indices

f16 = torch.float16


class _Site:
    def __init__(self, offset, nbytes, index):
        self.offset, self.nbytes, self.index, self.flag_offset = offset, nbytes, index, 0


class FakeArena:
    def __init__(self, rank, world, mems, barrier):
        self.rank, self.world, self.mems, self.barrier = rank, world, mems, barrier
        self.base = mems[rank].data_ptr()
        self.peer_base = [m.data_ptr() for m in mems]
        self.sites, self._cursor = {}, 256

    def site(self, key, nbytes):
        s = self.sites.get(key)
        if s is None:
            nbytes = (int(nbytes) + 255) // 256 * 256
            assert self._cursor + nbytes <= self.mems[self.rank].numel(), "fake arena too small"
            s = _Site(self._cursor, nbytes, len(self.sites))
            self._cursor += nbytes
            self.sites[key] = s
        return s

    def tensor(self, site, byte_offset, shape, dtype=f16):
        n = 1
        for d in shape:
            n *= int(d)
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert byte_offset + nb <= site.nbytes
        return self.mems[self.rank][site.offset + byte_offset: site.offset + byte_offset + nb].view(dtype).view(*shape)

    def peer_ptr(self, rank, site, byte_offset=0):
        return self.peer_base[rank] + site.offset + byte_offset

    def wait_mask(self, sources):
        return sum(1 << r for r in set(sources) if r != self.rank)

    def exchange(self, site, segs, sources, stream):
        for src, sp, r, dst, dp, rows, rb in segs:
            assert rb % 16 == 0 and sp % 16 == 0 and dp % 16 == 0 and src % 16 == 0 and dst % 16 == 0  # what fz_p2p_push requires
            lo, hi = self.peer_base[r] + site.offset, self.peer_base[r] + site.offset + site.nbytes
            assert lo <= dst and dst + (rows - 1) * dp + rb <= hi, "segment leaves the destination site"
            for i in range(rows):
                ctypes.memmove(dst + i * dp, src + i * sp, rb)
        self.barrier.wait()  # every rank has pushed: stronger than the per-source flags, equivalent for the data


def _run_ranks(world, fn):
    """fn(rank, engine) in one thread per rank; returns the per-rank results."""
    mems = [torch.zeros(8 << 20, dtype=torch.uint8) for _ in range(world)]
    barrier = threading.Barrier(world)
    out, err = [None] * world, []

    def work(r):
        try:
            e = object.__new__(UNetEngine)
            e.shard, e.dev, e.arena = (r, world, None), torch.device("cpu"), FakeArena(r, world, mems, barrier)
            out[r] = fn(r, e)
        except BaseException as ex:  # noqa: BLE001
            err.append(ex)
            barrier.abort()
    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if err:
        raise err[0]
    return out


@pytest.fixture(autouse=True)
def _no_cuda_stream(monkeypatch):
    monkeypatch.setattr(ops, "_stream", lambda: None)


@pytest.mark.parametrize("world,F", [(2, 4), (4, 2), (8, 1)])
def test_halo_extension(world, F):
    B, HW, C = 2, 8, 16
    y = torch.randn(B, world * F, HW, C).half()

    def fn(r, e):
        mine = y[:, r * F:(r + 1) * F].contiguous()
        return e._halo_ext(("halo_y", "layer"), mine), mine
    res = _run_ranks(world, fn)
    for r, (ext, mine) in enumerate(res):
        assert ext.shape == (B, F + 2, HW, C)
        assert torch.equal(ext[:, 1:F + 1], mine)
        left = y[:, r * F - 1] if r > 0 else torch.zeros(B, HW, C).half()          # zero padding of the un-sharded Conv1d at the clip ends
        right = y[:, (r + 1) * F] if r < world - 1 else torch.zeros(B, HW, C).half()
        assert torch.equal(ext[:, 0], left) and torch.equal(ext[:, F + 1], right)


@pytest.mark.parametrize("world,F,index_list", [(2, 2, ["mid"]), (4, 2, ["mid", "first"]), (8, 1, ["last"]), (2, 3, [-1, "first"]), (4, 1, ["mid", 1])])
def test_kv_exchange(world, F, index_list):
    B, S, heads, d = 2, 8, 2, 8
    C = heads * d
    Ft = world * F
    K = torch.randn(B, Ft, S, C).half()
    V = torch.randn(B, Ft, S, C).half()
    fis = sc_frame_indices(index_list, Ft)

    def fn(r, e):
        qk = torch.zeros(B * F * S, 2 * C, dtype=f16)
        qk[:, C:] = K[:, r * F:(r + 1) * F].reshape(B * F * S, C)
        vt = V[:, r * F:(r + 1) * F].reshape(B * F, S, heads, d).permute(0, 2, 3, 1).contiguous()
        return e._kv_exchange("layer", qk, vt, index_list, B, F, S, C, heads, d)
    res = _run_ranks(world, fn)
    for r, (k_src, vt_src, n_src, src_index) in enumerate(res):
        assert len(src_index) == len(index_list) and k_src.shape == (n_src * S, C) and vt_src.shape == (n_src, heads, d, S)
        for sl, fi in enumerate(fis):
            for b in range(B):
                for f in range(F):
                    row = src_index[sl][b * F + f]                                  # which block of the source buffer this query frame reads
                    g = fi[r * F + f]                                              # the GLOBAL frame attention_register.py:168-188 selects
                    assert torch.equal(k_src[row * S:(row + 1) * S], K[b, g]), (r, sl, b, f)
                    assert torch.equal(vt_src[row], V[b, g].reshape(S, heads, d).permute(1, 2, 0)), (r, sl, b, f)


def _temporal_attn_cpu(qkv, B, F, HW, heads, d, scale):
    """torch restatement of fz_temporal_attn_f16: qkv [B*F*HW, 3C] -> [B*F*HW, C], attention over the F frames of every (b, pixel, head)."""
    C = heads * d
    x = qkv.float().view(B, F, HW, 3, heads, d)
    q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]                         # [B, F, HW, heads, d]
    s = torch.einsum("bfphd,bgphd->bphfg", q, k) * scale
    p = torch.softmax(s, -1)
    o = torch.einsum("bphfg,bgphd->bfphd", p, v)
    return o.reshape(B * F * HW, C).half()


@pytest.mark.parametrize("world,F", [(2, 2), (4, 1), (8, 1)])
def test_temporal_attention_frames_pixels_exchange(world, F, monkeypatch):
    monkeypatch.setattr(ops, "temporal_attn", _temporal_attn_cpu)
    B, S, heads, d = 2, 16, 2, 8
    C = heads * d
    Ft = world * F
    qkv = torch.randn(B, Ft, S, 3 * C).half()
    want = _temporal_attn_cpu(qkv.reshape(B * Ft * S, 3 * C), B, Ft, S, heads, d, d ** -0.5).view(B, Ft, S, C)

    def fn(r, e):
        mine = qkv[:, r * F:(r + 1) * F].reshape(B * F * S, 3 * C).contiguous()
        return e._temporal_attn_sharded("layer", mine, B, F, S, heads, d, d ** -0.5).clone()
    res = _run_ranks(world, fn)
    for r, got in enumerate(res):
        assert torch.equal(got.view(B, F, S, C), want[:, r * F:(r + 1) * F]), r


@pytest.mark.parametrize("world,F", [(2, 2), (8, 1)])
def test_conv_out_gather(world, F, monkeypatch):
    """_finish gathers the 16-channel conv_out rows of ALL frames before the temporal tail and keeps the rank's frames."""
    B, H, W, co = 2, 4, 4, 4
    Ft = world * F
    y = torch.randn(B, Ft, H * W, 16).half()
    seen = {}

    def fake_out_temporal(yy, B_, co_, F_, H_, W_, **kw):
        seen[threading.get_ident()] = yy.clone()
        return yy.float().view(B_, F_, H_ * W_, 16)[..., :co_].permute(0, 3, 1, 2).reshape(B_, co_, F_, H_, W_).contiguous()
    monkeypatch.setattr(ops, "out_temporal", fake_out_temporal)

    def fn(r, e):
        e.w = {"conv_out.conv_temporal.down.weight": None, "conv_out.conv_temporal.down.weight#f32": torch.zeros(1),
               "conv_out.conv_temporal.up.weight#f32": torch.zeros(1)}
        e.lora_skip = {}
        mine = y[:, r * F:(r + 1) * F].reshape(B * F * H * W, 16).contiguous()
        return e._finish(mine, B, co, F, H, W)
    res = _run_ranks(world, fn)
    full = y.float()[..., :co].permute(0, 3, 1, 2).reshape(B, co, Ft, H, W)
    for r, got in enumerate(res):
        assert torch.equal(got, full[:, :, r * F:(r + 1) * F]), r
    for g in seen.values():
        assert torch.equal(g.view(B, Ft, H * W, 16), y)
