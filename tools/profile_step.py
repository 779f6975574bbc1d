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
pipe = bench.build_pipe(dev)
pipe.scheduler.set_timesteps(N)
from fatezero_b200 import synth  # noqa: E402
x0 = (synth.synth_latents(bench.FRAMES, bench.SIZE, bench.SIZE) * 0.5).to(dev)
emb = pipe._encode_prompt(bench.SRC, dev, 1, True, None)


def run(profile: bool):
    pipe.scheduler.set_timesteps(N)
    pipe.store_controller = controllers.AttentionStore()
    controllers.register_attention_control(pipe, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = True
    if profile:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    inv = pipe.ddim_clean2noisy_loop(x0, emb, pipe.store_controller)
    pipe.store_controller.LOW_RESOURCE = False
    pipe(prompt=bench.TGT, source_prompt=bench.SRC, edit_type="swap", image=None, strength=None, generator=None, num_inference_steps=N,
         clip_length=bench.FRAMES, guidance_scale=7.5, num_images_per_prompt=1, latents=inv[-1], uncond_embeddings_list=None,
         save_path=None, height=512, width=512, output_type="latent", use_inversion_attention=True, save_self_attention=False, **bench.P2P)
    torch.cuda.synchronize()
    if profile:
        torch.cuda.profiler.stop()


run(False)
run(True)
print("profiled", N, "inversion +", N, "edit steps")
