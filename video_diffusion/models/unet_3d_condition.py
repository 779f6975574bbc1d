from fatezero_b200.unet import UNetPseudo3DConditionModel, UNetPseudo3DConditionOutput  # noqa: F401
