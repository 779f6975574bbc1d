"""In-situ cost of each kernel class: time one full clip edit (bench workload) with that class's launches skipped (FZ_ABLATE).
The wall-clock delta to the full run includes launch gaps, cache state and tail effects that isolated micro-benchmarks miss;
"everything skipped" is the host-side floor (Python + ctypes + tensor-map encoding).  Results of ablated runs are garbage by design.
Usage (GPU box): python tools/ablate.py > gpurun_out/ablate.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = {
    "none": "",
    "gemm": "fz_gemm_f16",
    "conv3x3": "fz_conv3x3_nhwc_f16",
    "tconv3": "fz_tconv3_f16",
    "attention": "fz_attention_f16",
    "groupnorm": "fz_groupnorm_nhwc_f16",
    "layernorm": "fz_layernorm_f16",
    "temporal_attn": "fz_temporal_attn_f16",
    "misc": "fz_upsample2x_nhwc_f16,fz_concat_channels_f16,fz_im2col_latents_f16,fz_out_temporal_f32,fz_rowvec_linear,fz_timestep_sinusoid,"
            "fz_ddim_invert_step,fz_cfg_ddim_step,fz_blend_mask",
}
CLASSES["all"] = ",".join(v for k, v in CLASSES.items() if v)
out = {}
for name, skip in CLASSES.items():
    env = dict(os.environ, FZ_ABLATE=skip)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], env=env,
                       capture_output=True, text=True, timeout=600)
    try:
        line = json.loads(r.stdout.strip().splitlines()[-1])
        out[name] = line["ms_per_step"]
    except Exception as e:  # noqa: BLE001
        out[name] = f"failed: {e}: {r.stderr[-300:]}"
    print(name, out[name], file=sys.stderr, flush=True)
full = out.get("none")
table = {k: (round(full - v, 1) if isinstance(v, (int, float)) and isinstance(full, (int, float)) else v) for k, v in out.items()}
print(json.dumps({"ms_per_clip": out, "delta_ms_vs_full": table}, indent=1))
