"""TEST INFRASTRUCTURE ONLY (tier rule ③): nothing under fatezero_b200/ or video_diffusion/ may import this package.

oracle/shim/        tests-only restatement of the diffusers-0.11.1 symbols the reference imports (+ omegaconf/imageio stubs)
oracle/ref_harness  runs the UNMODIFIED reference (/root/reference, build container only) through the shim
oracle/fz_oracle    portable CPU fp32 restatement of the hot path (UNet forward, attention hooks, controllers, DDIM loops)
oracle/make_golden  writes tests/golden/*.pt from the reference (the pin for fz_oracle and for the CUDA path)
"""
