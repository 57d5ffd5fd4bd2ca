"""Input side of the training path: COCO-format datasets, the LSNet train pipelines and batching under the
reference's DATASETS / PIPELINES names, plus the GPU image preparation (``DevicePrep``)."""
from .coco import CocoDataset, CocoIndex, CocoPoseDataset  # noqa: F401
from .contour import PolygonMasks, unify_polygons, uniformsample  # noqa: F401
from .loader import (DevicePrep, DistributedGroupSampler, GroupSampler, build_dataloader, build_dataset,  # noqa: F401
                     collate, device_prep_pipeline)
from .transforms import (Collect, Compose, DefaultFormatBundle, DeviceFormatBundle, ImageToTensor,  # noqa: F401
                         LoadAnnotations, LoadImageFromFile, MultiScaleFlipAug, Normalize, Pad, RandomFlip, Resize)
