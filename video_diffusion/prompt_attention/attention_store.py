from fatezero_b200.controllers import AttentionControl, AttentionStore  # noqa: F401
