"""x3d_tf_b200 -- B200-native (sm_100a) implementation of the X3D forward path.

Same class API as the reference's `model.py`; all arithmetic runs in hand-written CUDA kernels
behind the C ABI of `include/x3d_b200.h` (see DESIGN.md / INTEGRATION.md).
"""
from .config import CfgNode, get_config, get_default_config, three_crop_size  # noqa: F401

__all__ = ["CfgNode", "get_config", "get_default_config", "three_crop_size", "X3D", "X3D_Stem",
           "Bottleneck", "ResBlock", "ResStage", "AdaptiveAvgPool3D", "reset_block_counters"]


def __getattr__(name):            # the model module needs torch; keep config-only imports light
    if name in ("X3D", "X3D_Stem", "Bottleneck", "ResBlock", "ResStage", "AdaptiveAvgPool3D",
                "reset_block_counters", "Options"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
