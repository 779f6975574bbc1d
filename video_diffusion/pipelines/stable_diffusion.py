from fatezero_b200.pipeline import SpatioTemporalStableDiffusionPipeline  # noqa: F401
