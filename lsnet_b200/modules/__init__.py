from .assigners import ATSSAssigner, AssignResult, CentroidAssigner, PseudoSampler  # noqa: F401
from .backbone import ResNet, ResNeXt  # noqa: F401
from .dcn import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack,  # noqa: F401
                  PyramidDeformConv)
from .detector import LSDetector  # noqa: F401
from .fpn import FPN  # noqa: F401
from .head import DCNConvModule, LSHead  # noqa: F401
from .losses import CrossIOULoss, FocalLoss  # noqa: F401
