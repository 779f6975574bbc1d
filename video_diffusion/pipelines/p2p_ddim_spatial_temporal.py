from fatezero_b200.pipeline import P2pDDIMSpatioTemporalPipeline, StableDiffusionPipelineOutput  # noqa: F401
