"""lsnet_b200 — B200-native (sm_100a) LSNet training hot path behind the reference's registry / config surface.
Importing the package registers LSDetector, ResNet/ResNeXt, FPN, LSHead, CrossIOULoss, FocalLoss, CentroidAssigner,
ATSSAssigner, DCN/DCNv2 and the input side (CocoDataset, CocoPoseDataset, the train pipelines) in lsnet_b200.registry;
nothing here touches oracle/."""
from . import datasets, modules  # noqa: F401
from .config import Config  # noqa: F401
from .registry import (BACKBONES, BBOX_ASSIGNERS, CONV_LAYERS, DATASETS, DETECTORS, HEADS, LOSSES, NECKS,  # noqa: F401
                       PIPELINES,
                       build_assigner, build_backbone, build_detector, build_head, build_loss, build_neck)

__version__ = '0.1.0'
