"""Latency / bandwidth of the frame-shard exchange primitive (fz_p2p.cu) between the ranks of this job.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_p2p.py"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FZ_P2P_ARENA_MB", "1024")
from fatezero_b200 import _lib, dist as fzdist, ops, p2p  # noqa: E402


def main():
    rank, world, local = fzdist.env_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    fzdist.init("nccl", dev)
    _lib.check(_lib.load().fz_init(ops._stream()), "fz_init")
    ar = p2p.Arena(rank, world, dev)
    st = ops._stream()
    out = {}
    left, right = (rank - 1) % world, (rank + 1) % world
    src = torch.randn(64 << 20, device=dev).half()  # 128 MiB source
    for name, halo_bytes, local_bytes in (("flag_only_16B", 16, 0), ("halo_64KB", 64 << 10, 0), ("halo_2.6MB", 2621440, 0),
                                          ("halo_2.6MB+local_21MB", 2621440, 20971520), ("local_21MB_only", 0, 20971520)):
        site = ar.site(("bench", name), 2 * max(halo_bytes, 16) + max(local_bytes, 16) + 4096)
        segs, srcs = [], []
        if halo_bytes:
            segs.append((src.data_ptr(), halo_bytes, right, ar.peer_ptr(right, site, 0), halo_bytes, 1, halo_bytes))
            srcs.append(left)
            if world > 2:
                segs.append((src.data_ptr() + halo_bytes, halo_bytes, left, ar.peer_ptr(left, site, halo_bytes), halo_bytes, 1, halo_bytes))
                srcs.append(right)
        if local_bytes:
            segs.append((src.data_ptr(), local_bytes, rank, ar.peer_ptr(rank, site, 2 * max(halo_bytes, 16)), local_bytes, 1, local_bytes))
        for reps in (20, 200):
            dist.barrier()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                ar.exchange(site, segs, srcs, st)
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) * 1e3 / reps
        out[name] = round(us, 2)
    # the same launches inside a CUDA graph (what the product replays)
    site = ar.site(("bench", "graph"), 4096)
    segs = [(src.data_ptr(), 64 << 10, right, ar.peer_ptr(right, site, 0), 64 << 10, 1, 64 << 10)] if world > 1 else []
    site = ar.site(("bench", "graph2"), (64 << 10) + 4096)
    segs = [(src.data_ptr(), 64 << 10, right, ar.peer_ptr(right, site, 0), 64 << 10, 1, 64 << 10)]
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g.capture_begin(capture_error_mode="thread_local")
        for _ in range(100):
            ar.exchange(site, segs, [left], ops._stream())
        g.capture_end()
    torch.cuda.current_stream().wait_stream(side)
    for reps in (2, 10):
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / (reps * 100)
    out["graph_halo_64KB"] = round(us, 2)
    # plain small kernel launch floor for reference: 100 LayerNorms of a tiny tensor in a graph
    x = torch.randn(256, 320, device=dev).half()
    gm, bt = torch.ones(320, device=dev), torch.zeros(320, device=dev)
    g2 = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g2.capture_begin(capture_error_mode="thread_local")
        for _ in range(100):
            ops.layernorm(x, gm, bt)
        g2.capture_end()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        g2.replay()
    e.record()
    torch.cuda.synchronize()
    out["graph_tiny_layernorm_kernel"] = round(s.elapsed_time(e) * 1e3 / 1000, 2)
    out.update(world=world, unit="us per launch (max over nothing: this rank)")
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
