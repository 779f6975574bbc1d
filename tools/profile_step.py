"""One inversion DDIM step (STORE) + one edit DDIM step (INJECT) of the bench workload, bracketed by cudaProfilerStart/Stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tapgemm -c 3 -o gpurun_out/prof_gemm \
      python tools/profile_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fatezero_b200 import controllers  # noqa: E402

N = int(os.environ.get("FZ_PROFILE_STEPS", "2"))
dev = torch.device("cuda", 0)
bench.select_config(os.environ.get("FZ_PROFILE_CONFIG", "style"))
CFG = bench.CFG
pipe = bench.build_pipe(dev)
pipe.graph_mode = "off"  # per-kernel launches for the profiler
pipe.scheduler.set_timesteps(N)
from fatezero_b200 import synth  # noqa: E402
x0 = (synth.synth_latents(CFG["frames"], CFG["size"], CFG["size"]) * 0.5).to(dev)
emb = pipe._encode_prompt(bench.SRC, dev, 1, True, None)


SEQ = []  # (entry point, shape key) of every tensor-core launch of the profiled pass, in launch order (joined with the ncu list by index)


def _record_shapes():
    from fatezero_b200 import ops

    def wrap(name, fn, key):
        def inner(*a, **k):
            SEQ.append((name, key(a, k)))
            return fn(*a, **k)
        return inner
    ops.gemm = wrap("gemm", ops.gemm, lambda a, k: [list(a[0].shape), list(a[1].shape), bool(k.get("geglu")), k.get("residual") is not None])
    ops.conv3x3 = wrap("conv3x3", ops.conv3x3, lambda a, k: [list(a[0].shape), list(a[1].shape), k.get("stride", 1)])
    ops.tconv3 = wrap("tconv3", ops.tconv3, lambda a, k: [list(a[0].shape), list(a[1].shape)])
    ops.attention = wrap("attention", ops.attention, lambda a, k: [k["S_q"], len(k["src_index"]) * k["keys_per_slot"], k["d"], k["BF"], k.get("row_mode", 0),
                                                                  k.get("edit_bf_start", 0)])


def run(profile: bool):
    pipe.scheduler.set_timesteps(N)
    pipe.store_controller = controllers.AttentionStore()
    controllers.register_attention_control(pipe, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = True
    if profile:
        _record_shapes()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    inv = pipe.ddim_clean2noisy_loop(x0, emb, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = False
    import tempfile
    pipe(prompt=CFG["tgt"], source_prompt=bench.SRC, edit_type="swap", image=None, strength=None, generator=None, num_inference_steps=N,
         clip_length=CFG["frames"], guidance_scale=7.5, num_images_per_prompt=1, latents=inv[-1], uncond_embeddings_list=None,
         save_path=tempfile.mkdtemp() if CFG["p2p"].get("blend_words") else None, height=8 * CFG["size"], width=8 * CFG["size"],
         output_type="latent", use_inversion_attention=True, save_self_attention=False, **CFG["p2p"])
    torch.cuda.synchronize()
    if profile:
        torch.cuda.profiler.stop()


run(False)
run(True)
import json  # noqa: E402
with open(os.path.join(ROOT, "gpurun_out", "profile_step_shapes.json"), "w") as f:
    json.dump(SEQ, f)
print("profiled", N, "inversion +", N, "edit steps;", len(SEQ), "tensor-core launches recorded")
