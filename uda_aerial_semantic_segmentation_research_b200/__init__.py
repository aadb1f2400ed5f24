"""uda_aerial_semantic_segmentation_research_b200 — B200-native (sm_100a) training hot path for
bempt/uda_aerial_semantic_segmentation_research, behind the reference's own entry points.

    from uda_aerial_semantic_segmentation_research_b200 import Unet, DiceLoss, ...

See DESIGN.md (kernels, layout, rooflines) and INTEGRATION.md (how the reference binds it).
"""
from .unet import Unet, create_model  # noqa: F401

__all__ = ["Unet", "create_model"]
__version__ = "0.1.0"
