"""drn_wsod_pytorch_b200 -- B200-native (sm_100a) implementation of the DRN-WSOD per-image
detection hot path behind the reference's registry / config / state_dict surface.

    from drn_wsod_pytorch_b200 import builtin_config, build_model
    model = build_model(builtin_config("oicr_WSR_18_DC5_1x"))      # needs libdrn_b200.so + a B200
"""
from . import data  # noqa: F401  (proposal files, transform_proposals, weight files)
from .config import CfgNode, builtin_config, get_cfg  # noqa: F401
from .modeling import (  # noqa: F401
    DiscriminativeAdaptionNeck,
    GeneralizedRCNNWSL,
    OICRROIHeads,
    PCLROIHeads,
    WSDDNROIHeads,
    build_model,
    build_vgg_backbone,
    build_ws_resnet_backbone,
)
from .registry import (  # noqa: F401
    BACKBONE_REGISTRY,
    META_ARCH_REGISTRY,
    ROI_BOX_HEAD_REGISTRY,
    ROI_HEADS_REGISTRY,
    register_into_detectron2,
)
from .solver import FusedSGD, build_optimizer, run_step  # noqa: F401
from .structures import Boxes, ImageList, Instances  # noqa: F401
from .tta import DatasetMapperTTAAVG, DatasetMapperTTAUNION, GeneralizedRCNNWithTTAAVG, GeneralizedRCNNWithTTAUNION  # noqa: F401

__version__ = "0.1.0"
