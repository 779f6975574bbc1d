from fatezero_b200.spatial_blend import SpatialBlender  # noqa: F401
