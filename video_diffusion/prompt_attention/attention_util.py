from fatezero_b200.controllers import (AttentionControl, AttentionControlEdit, AttentionRefine, AttentionReplace,  # noqa: F401
                                       AttentionReweight, AttentionStore, EmptyControl, get_equalizer, make_controller,
                                       register_attention_control)
from fatezero_b200.spatial_blend import SpatialBlender  # noqa: F401
from fatezero_b200.visualization import show_cross_attention  # noqa: F401
