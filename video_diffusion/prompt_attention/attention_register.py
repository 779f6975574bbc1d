from fatezero_b200.controllers import register_attention_control  # noqa: F401
