from fatezero_b200.visualization import aggregate_attention, show_cross_attention  # noqa: F401
