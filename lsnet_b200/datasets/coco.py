"""COCO-format datasets for the three LSNet tasks under the reference's DATASETS names.

Reference behaviour: mmdet/datasets/custom.py:12-213 (the dataset protocol: ``pre_pipeline`` field lists, aspect-ratio
group flags, redraw on a rejected sample), coco.py:32-186 (``CocoDataset``: detection / contour annotations with
pre-computed ``extreme_points``) and coco_pose.py:26-170 (``CocoPoseDataset``: person keypoints).  The annotation json
is indexed directly (a dict of lists) — pycocotools is not a dependency.
"""
import json
from collections import defaultdict

import numpy as np

from ..registry import DATASETS
from .transforms import Compose

COCO_CLASSES = (
    'person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light',
    'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant',
    'bear', 'zebra', 'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard',
    'sports ball', 'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket', 'bottle',
    'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli',
    'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair', 'couch', 'potted plant', 'bed', 'dining table', 'toilet',
    'tv', 'laptop', 'mouse', 'remote', 'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink',
    'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear', 'hair drier', 'toothbrush')


class CocoIndex:
    """The three lookups the datasets need from a COCO json (what ``pycocotools.coco.COCO`` provides through
    get_cat_ids / get_img_ids / load_imgs / get_ann_ids / load_anns)."""

    def __init__(self, ann_file):
        if isinstance(ann_file, dict):
            data = ann_file
        else:
            with open(ann_file) as f:
                data = json.load(f)
        self.cats = {c['id']: c for c in data.get('categories', [])}
        self.imgs = {im['id']: im for im in data['images']}
        self.img_anns = defaultdict(list)
        for a in data.get('annotations', []):
            self.img_anns[a['image_id']].append(a)

    def cat_ids(self, names=None):
        if names is None:
            return list(self.cats)
        # one id per requested name, in the json's category order (pycocotools getCatIds(catNms=…))
        return [c['id'] for c in self.cats.values() if c['name'] in set(names)]

    def img_ids(self):
        return list(self.imgs)                 # json order, as pycocotools' getImgIds()


@DATASETS.register_module()
class CocoDataset:
    CLASSES = COCO_CLASSES
    #: key of ``ann_info`` -> (annotation field, row width) of the per-instance landmark table this dataset adds
    LANDMARK_FIELD = ('extremes', 'extreme_points', 10)

    def __init__(self, ann_file, pipeline, classes=None, data_root=None, img_prefix='', seg_prefix=None,
                 proposal_file=None, test_mode=False, filter_empty_gt=True):
        import os.path as osp
        if data_root is not None:
            if isinstance(ann_file, str) and not osp.isabs(ann_file):
                ann_file = osp.join(data_root, ann_file)
            if img_prefix and not osp.isabs(img_prefix):
                img_prefix = osp.join(data_root, img_prefix)
        self.ann_file, self.img_prefix, self.seg_prefix, self.proposal_file = ann_file, img_prefix, seg_prefix, None
        self.test_mode, self.filter_empty_gt = test_mode, filter_empty_gt
        self.CLASSES = self.get_classes(classes)
        self.data_infos = self.load_annotations(ann_file)
        if self.custom_classes:
            self.data_infos = self.get_subset_by_classes()
        if not test_mode:
            keep = self._filter_imgs()
            self.data_infos = [self.data_infos[i] for i in keep]
            self.img_ids = [self.img_ids[i] for i in keep]
            self._set_group_flag()
        self.pipeline = Compose(pipeline)

    def get_classes(self, classes=None):
        """custom.py:233-257: None keeps the dataset's CLASSES; a list / tuple overrides them (kept as given); a string
        is a file with one class name per line."""
        self.custom_classes = classes is not None
        if classes is None:
            return type(self).CLASSES
        if isinstance(classes, str):
            with open(classes) as f:
                return [line.rstrip('\n\r') for line in f if line.strip()]
        if isinstance(classes, (tuple, list)):
            return classes
        raise ValueError(f'Unsupported type {type(classes)} of classes.')

    def get_subset_by_classes(self):
        """coco.py:98-121: with custom classes only the images holding at least one instance of them are kept (json
        order here; the reference iterates a set)."""
        wanted = set(self.cat_ids)
        keep = {i for i, anns in self.coco.img_anns.items() if any(a['category_id'] in wanted for a in anns)}
        self.img_ids = [i for i in self.img_ids if i in keep]
        infos = []
        for i in self.img_ids:
            info = dict(self.coco.imgs[i])
            info['filename'] = info['file_name']
            infos.append(info)
        return infos

    # -- coco.py:32-60 --------------------------------------------------------------------------------------------
    def load_annotations(self, ann_file):
        self.coco = CocoIndex(ann_file)
        self.cat_ids = self.coco.cat_ids(self.CLASSES)
        self.cat2label = {cid: i for i, cid in enumerate(self.cat_ids)}
        self.img_ids = self.coco.img_ids()
        infos = []
        for i in self.img_ids:
            info = dict(self.coco.imgs[i])
            info['filename'] = info['file_name']
            infos.append(info)
        return infos

    def __len__(self):
        return len(self.data_infos)

    def get_ann_info(self, idx):
        return self._parse_ann_info(self.data_infos[idx], self.coco.img_anns.get(self.data_infos[idx]['id'], []))

    def get_cat_ids(self, idx):
        return [a['category_id'] for a in self.coco.img_anns.get(self.data_infos[idx]['id'], [])]

    def _filter_imgs(self, min_size=32):
        """coco.py:95-104: drop images smaller than ``min_size`` and (filter_empty_gt) images without annotations."""
        with_ann = set(self.coco.img_anns)
        keep = []
        for i, info in enumerate(self.data_infos):
            if self.filter_empty_gt and self.img_ids[i] not in with_ann:
                continue
            if min(info['width'], info['height']) >= min_size:
                keep.append(i)
        return keep

    def _set_group_flag(self):
        """custom.py:158-168: group 1 = landscape images; a batch never mixes the two groups."""
        self.flag = np.array([1 if info['width'] / info['height'] > 1 else 0 for info in self.data_infos], np.uint8)

    def _keep(self, ann, img_info):
        """The instance filter shared by both datasets (coco.py:141-151): ignored, outside the image, degenerate or
        of a class that is not trained."""
        if ann.get('ignore', False):
            return False
        x1, y1, w, h = ann['bbox']
        iw = max(0, min(x1 + w, img_info['width']) - max(x1, 0))
        ih = max(0, min(y1 + h, img_info['height']) - max(y1, 0))
        if iw * ih == 0 or ann['area'] <= 0 or w < 1 or h < 1:
            return False
        return ann['category_id'] in self.cat2label

    def _parse_ann_info(self, img_info, anns):
        key, field, width = self.LANDMARK_FIELD
        boxes, labels, ignore, masks, marks = [], [], [], [], []
        for a in anns:
            if not self._keep(a, img_info):
                continue
            x1, y1, w, h = a['bbox']
            box = [x1, y1, x1 + w, y1 + h]
            if a.get('iscrowd', False):
                ignore.append(box)
            else:
                boxes.append(box)
                labels.append(self.cat2label[a['category_id']])
                masks.append(a.get('segmentation'))
                marks.append(a[field] if field in a else None)
        if marks and any(m is None for m in marks):
            if all(m is None for m in marks):
                marks = None                       # annotation file without this landmark type (e.g. plain instances)
            else:
                raise KeyError(f'some annotations of image {img_info["id"]} lack "{field}"')
        out = dict(bboxes=np.array(boxes, np.float32).reshape(-1, 4), labels=np.array(labels, np.int64),
                   bboxes_ignore=np.array(ignore, np.float32).reshape(-1, 4), masks=masks,
                   seg_map=img_info['filename'].replace('jpg', 'png'))
        if marks is not None:
            out[key] = np.array(marks, np.float32).reshape(-1, width)
        return out

    # -- custom.py:139-213 ------------------------------------------------------------------------------------------
    def pre_pipeline(self, results):
        results['img_prefix'] = self.img_prefix
        results['seg_prefix'] = self.seg_prefix
        results['proposal_file'] = self.proposal_file
        for f in ('bbox_fields', 'extreme_fields', 'mask_fields', 'seg_fields', 'keypoint_fields'):
            results[f] = []

    def _rand_another(self, idx):
        return int(np.random.choice(np.where(self.flag == self.flag[idx])[0]))

    def prepare_train_img(self, idx):
        results = dict(img_info=self.data_infos[idx], ann_info=self.get_ann_info(idx))
        self.pre_pipeline(results)
        return self.pipeline(results)

    def prepare_test_img(self, idx):
        results = dict(img_info=self.data_infos[idx])
        self.pre_pipeline(results)
        return self.pipeline(results)

    def __getitem__(self, idx):
        if self.test_mode:
            return self.prepare_test_img(idx)
        while True:
            data = self.prepare_train_img(idx)
            if data is None or len(data.get('gt_labels', [0])) == 0:
                # the reference head cannot take an image without instances either (torch.stack of an empty list,
                # lsnet_head.py:1737); filter_empty_gt keeps such images out unless every instance was filtered
                idx = self._rand_another(idx)
                continue
            return data


    # -- results -> COCO json (coco_pose.py:174-329; LSNet results are [boxes per class, landmark vectors per class]) --
    @staticmethod
    def xyxy2xywh(bbox):
        b = bbox.tolist()
        return [b[0], b[1], b[2] - b[0], b[3] - b[1]]

    def _records(self, results):
        """(image id, category id, box row (5,), vector row) of every detection, in the reference's nesting order
        (image, class, detection)."""
        for idx in range(len(self)):
            boxes, vecs = results[idx][0], results[idx][1]
            for label in range(len(boxes)):
                for i in range(boxes[label].shape[0]):
                    yield self.img_ids[idx], self.cat_ids[label], boxes[label][i], vecs[label][i]

    def _det2json(self, results):
        return [dict(image_id=im, bbox=self.xyxy2xywh(b), score=float(b[4]), category_id=c)
                for im, c, b, _ in self._records(results)]

    def _poly2json(self, results):
        """Contour results as COCO polygon segmentations (one ring of ``num_vectors`` points per instance); mask AP then
        comes from pycocotools, which rasterises polygons itself."""
        return [dict(image_id=im, bbox=self.xyxy2xywh(b), score=float(b[4]), category_id=c,
                     segmentation=[[float(x) for x in v]]) for im, c, b, v in self._records(results)]

    def results2json(self, results, outfile_prefix):
        """``<prefix>.bbox.json`` always; ``<prefix>.segm.json`` when the vectors are contours (more than the 4 extreme
        points).  Returns the dict of written files (keys as coco.py:281-320)."""
        import json as _json
        files = dict(bbox=f'{outfile_prefix}.bbox.json', proposal=f'{outfile_prefix}.bbox.json')
        with open(files['bbox'], 'w') as f:
            _json.dump(self._det2json(results), f)
        width = next((v.shape[1] for r in results for v in r[1] if v.shape[0]), 0)
        if width > 8:
            files['segm'] = f'{outfile_prefix}.segm.json'
            with open(files['segm'], 'w') as f:
                _json.dump(self._poly2json(results), f)
        return files

    def format_results(self, results, jsonfile_prefix=None, **kwargs):
        """coco.py:342-368."""
        import os.path as osp
        import tempfile
        assert isinstance(results, list), 'results must be a list'
        assert len(results) == len(self), f'The length of results is not equal to the dataset len: {len(results)} != {len(self)}'
        tmp = None
        if jsonfile_prefix is None:
            tmp = tempfile.TemporaryDirectory()
            jsonfile_prefix = osp.join(tmp.name, 'results')
        return self.results2json(results, jsonfile_prefix), tmp

    def evaluate(self, results, metric='bbox', **kwargs):
        raise NotImplementedError('COCO AP needs pycocotools (COCOeval), which is outside this path: write the json files '
                                  'with format_results() and evaluate them with the COCO API')


@DATASETS.register_module()
class CocoPoseDataset(CocoDataset):
    """coco_pose.py:26-170: person keypoints, rows of 17 × [x, y, v]."""
    CLASSES = ('person',)
    LANDMARK_FIELD = ('keypoints', 'keypoints', 51)

    def _kps2json(self, results):
        """coco_pose.py:226-247: 17 (x, y) landmarks -> COCO keypoints with visibility 1, scored by the box score."""
        out = []
        for im, c, b, v in self._records(results):
            kps = np.concatenate([v.reshape(-1, 2), np.ones((17, 1), dtype=np.float32)], axis=1).reshape(51).tolist()
            out.append(dict(image_id=im, bbox=self.xyxy2xywh(b), keypoints=kps, score=float(b[4]), category_id=c))
        return out

    def results2json(self, results, outfile_prefix):
        """coco_pose.py:287-329: ``.bbox.json`` and ``.kps.json``."""
        import json as _json
        files = dict(bbox=f'{outfile_prefix}.bbox.json', proposal=f'{outfile_prefix}.bbox.json',
                     keypoints=f'{outfile_prefix}.kps.json')
        with open(files['bbox'], 'w') as f:
            _json.dump(self._det2json(results), f)
        with open(files['keypoints'], 'w') as f:
            _json.dump(self._kps2json(results), f)
        return files
