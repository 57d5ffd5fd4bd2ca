"""The train and test pipelines of the LSNet configs, under the reference's PIPELINES names so
``cfg.data.train.pipeline`` / ``cfg.data.test.pipeline`` resolve unchanged:

    LoadImageFromFile -> LoadAnnotations(with_bbox, with_extreme | with_keypoint | with_mask+poly2mask=False)
    -> Resize(keep_ratio, single scale | multiscale range / value) -> RandomFlip -> Normalize -> Pad(size_divisor)
    -> DefaultFormatBundle -> Collect

(configs/_base_/datasets/coco_lsvr.py:5-14, coco_pose.py:5-14, configs/lsnet/lsnet_segm_r50_fpn_1x_coco.py:8-18).
Reference behaviour: mmdet/datasets/pipelines/{loading.py:11-107,183-470, transforms.py:25-570, formating.py:170-330,
compose.py}.  Every stage takes and returns the same ``results`` dict with the same keys, so the stages can be mixed
with reference ones.

(test: LoadImageFromFile -> MultiScaleFlipAug[Resize, RandomFlip, Normalize, Pad, ImageToTensor, Collect],
coco_lsvr.py:15-29.)

B200-first difference (opt-in, ``loader.device_prep_pipeline(pipeline)``): Normalize / Pad / the NCHW transpose are
left out of the worker pipeline; the batch carries the resized uint8 HWC images and ONE kernel
(``lsnet_image_prep_u8``, csrc/elementwise.cu) writes the normalised, zero-padded fp32 NCHW canvas on the GPU —
4× fewer bytes over PCIe per step and no float image passes on the host.
"""
import os.path as osp

import numpy as np
import torch

from ..registry import PIPELINES, build_from_cfg
from .contour import PolygonMasks, unify_polygons


def _cv2():
    import cv2
    return cv2


def rescale_size(old_size, scale):
    """mmcv/mmcv/image/geometric.py:6-17,76-110: (w, h) scaled by a factor, or as large as fits inside
    (max(scale), min(scale)); sizes round half up."""
    w, h = old_size
    if isinstance(scale, (float, int)):
        if scale <= 0:
            raise ValueError(f'Invalid scale {scale}, must be positive.')
        f = scale
    elif isinstance(scale, tuple):
        f = min(max(scale) / max(h, w), min(scale) / min(h, w))
    else:
        raise TypeError(f'Scale must be a number or tuple of int, but got {type(scale)}')
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


class Compose:
    """compose.py:8-51: a ``None`` from any stage aborts the sample (the dataset then draws another index)."""

    def __init__(self, transforms):
        self.transforms = []
        for t in transforms:
            if isinstance(t, dict):
                t = build_from_cfg(t, PIPELINES)
            elif not callable(t):
                raise TypeError('transform must be callable or a dict')
            self.transforms.append(t)

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
            if data is None:
                return None
        return data


@PIPELINES.register_module()
class LoadImageFromFile:
    """loading.py:11-107: BGR uint8 image from ``img_prefix / img_info['filename']`` (cv2 decode, as mmcv's default
    backend)."""

    def __init__(self, to_float32=False, color_type='color', file_client_args=None):
        self.to_float32 = to_float32
        self.color_type = color_type

    def __call__(self, results):
        name = results['img_info']['filename']
        path = osp.join(results['img_prefix'], name) if results.get('img_prefix') is not None else name
        cv2 = _cv2()
        img = cv2.imread(path, cv2.IMREAD_COLOR if self.color_type == 'color' else cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if self.to_float32:
            img = img.astype(np.float32)
        results['filename'] = path
        results['ori_filename'] = name
        results['img'] = img
        results['img_shape'] = img.shape
        results['ori_shape'] = img.shape
        results['img_fields'] = ['img']
        return results


@PIPELINES.register_module()
class LoadAnnotations:
    """loading.py:183-470.  ``poly2mask=True`` (bitmap masks through pycocotools) is not an LSNet path and raises."""

    def __init__(self, with_bbox=True, with_label=True, with_mask=False, with_seg=False, with_extreme=False,
                 with_keypoint=False, poly2mask=True, file_client_args=None, spline_num=10, num_contour_points=128):
        if with_seg:
            raise NotImplementedError('semantic segmentation maps are not on the LSNet path')
        if with_mask and poly2mask:
            raise NotImplementedError('LSNet trains on contours: use poly2mask=False (bitmap masks need pycocotools)')
        self.with_bbox, self.with_label, self.with_mask = with_bbox, with_label, with_mask
        self.with_extreme, self.with_keypoint = with_extreme, with_keypoint
        self.spline_num, self.num_points = spline_num, num_contour_points

    def __call__(self, results):
        ann = results['ann_info']
        if self.with_bbox:
            results['gt_bboxes'] = ann['bboxes'].copy()
            if ann.get('bboxes_ignore') is not None:
                results['gt_bboxes_ignore'] = ann['bboxes_ignore'].copy()
                results['bbox_fields'].append('gt_bboxes_ignore')
            results['bbox_fields'].append('gt_bboxes')
        if self.with_extreme:
            results['gt_extremes'] = ann['extremes'].copy()
            results['extreme_fields'].append('gt_extremes')
        if self.with_keypoint:
            results['gt_keypoints'] = ann['keypoints'].copy()
            results['keypoint_fields'].append('gt_keypoints')
        if self.with_label:
            results['gt_labels'] = ann['labels'].copy()
        if self.with_mask:
            h, w = results['img_info']['height'], results['img_info']['width']
            boxes = ann['bboxes']
            results['gt_masks'] = PolygonMasks(
                [unify_polygons(p, boxes[i], self.num_points, self.spline_num) for i, p in enumerate(ann['masks'])],
                h, w)
            results['mask_fields'].append('gt_masks')
        return results


@PIPELINES.register_module()
class Resize:
    """transforms.py:25-300: one scale, a uniformly drawn scale between two corner scales ('range'), one of a list
    ('value'), or ``ratio_range`` times a base scale; boxes / extreme points / keypoints are scaled and clipped to the
    new image, polygons scaled."""

    def __init__(self, img_scale=None, multiscale_mode='range', ratio_range=None, keep_ratio=True):
        if img_scale is None:
            self.img_scale = None
        else:
            self.img_scale = img_scale if isinstance(img_scale, list) else [img_scale]
            assert all(isinstance(s, tuple) for s in self.img_scale), 'img_scale: a tuple or a list of tuples'
        if ratio_range is not None:
            assert len(self.img_scale) == 1
        else:
            assert multiscale_mode in ['value', 'range']
        self.multiscale_mode, self.ratio_range, self.keep_ratio = multiscale_mode, ratio_range, keep_ratio

    @staticmethod
    def random_select(img_scales):
        i = np.random.randint(len(img_scales))
        return img_scales[i], i

    @staticmethod
    def random_sample(img_scales):
        """:98-122 — note the draw order (long edge first) matters for reproducing a seeded run."""
        assert len(img_scales) == 2
        longs = [max(s) for s in img_scales]
        shorts = [min(s) for s in img_scales]
        long_edge = np.random.randint(min(longs), max(longs) + 1)
        short_edge = np.random.randint(min(shorts), max(shorts) + 1)
        return (long_edge, short_edge), None

    @staticmethod
    def random_sample_ratio(img_scale, ratio_range):
        lo, hi = ratio_range
        assert lo <= hi
        r = np.random.random_sample() * (hi - lo) + lo
        return (int(img_scale[0] * r), int(img_scale[1] * r)), None

    def _random_scale(self, results):
        if self.ratio_range is not None:
            scale, idx = self.random_sample_ratio(self.img_scale[0], self.ratio_range)
        elif len(self.img_scale) == 1:
            scale, idx = self.img_scale[0], 0
        elif self.multiscale_mode == 'range':
            scale, idx = self.random_sample(self.img_scale)
        else:
            scale, idx = self.random_select(self.img_scale)
        results['scale'], results['scale_idx'] = scale, idx

    def __call__(self, results):
        if 'scale' not in results:
            if 'scale_factor' in results:
                f = results['scale_factor']
                assert isinstance(f, float)
                results['scale'] = tuple([int(x * f) for x in results['img'].shape[:2]][::-1])
            else:
                self._random_scale(results)
        else:
            assert 'scale_factor' not in results, 'scale and scale_factor cannot be both set.'
        cv2 = _cv2()
        for key in results.get('img_fields', ['img']):
            img = results[key]
            h, w = img.shape[:2]
            if self.keep_ratio:
                new_w, new_h = rescale_size((w, h), results['scale'])
            else:
                new_w, new_h = results['scale']
            out = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
            results[key] = out
            ws, hs = new_w / w, new_h / h
            results['img_shape'] = out.shape
            results['pad_shape'] = out.shape
            results['scale_factor'] = np.array([ws, hs, ws, hs], dtype=np.float32)
            results['keep_ratio'] = self.keep_ratio
        H, W = results['img_shape'][:2]
        sf = results['scale_factor']
        for key in results.get('bbox_fields', []):
            b = results[key] * sf
            b[:, 0::2] = np.clip(b[:, 0::2], 0, W)
            b[:, 1::2] = np.clip(b[:, 1::2], 0, H)
            results[key] = b
        for key in results.get('extreme_fields', []):
            e = results[key] * np.tile(sf[:2], (1, 5))
            e[:, 0::2] = np.clip(e[:, 0::2], 0, W)
            e[:, 1::2] = np.clip(e[:, 1::2], 0, H)
            results[key] = e
        for key in results.get('keypoint_fields', []):       # in place, as the reference (:228-239)
            k = results[key]
            k[:, 0::3] = np.clip(k[:, 0::3] * sf[0], 0, W)
            k[:, 1::3] = np.clip(k[:, 1::3] * sf[1], 0, H)
        for key in results.get('mask_fields', []):
            if results[key] is None:
                continue
            results[key] = results[key].rescale(results['scale']) if self.keep_ratio else \
                results[key].resize(results['img_shape'][:2])
        return results


#: COCO left/right keypoint pairs swapped by a horizontal flip (transforms.py:322-323)
KEYPOINT_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def bbox_flip(b, img_shape, direction):
    f = b.copy()
    if direction == 'horizontal':
        w = img_shape[1]
        f[..., 0::4] = w - b[..., 2::4]
        f[..., 2::4] = w - b[..., 0::4]
    elif direction == 'vertical':
        h = img_shape[0]
        f[..., 1::4] = h - b[..., 3::4]
        f[..., 3::4] = h - b[..., 1::4]
    else:
        raise ValueError(f"Invalid flipping direction '{direction}'")
    return f


def extreme_flip(e, img_shape, direction):
    """transforms.py:354-388.  Row layout: top, left, bottom, right, centre (x, y each).  A horizontal flip mirrors
    every x and swaps the left and right points; a vertical flip mirrors every y and swaps top and bottom."""
    f = e.copy()
    if direction == 'horizontal':
        w = img_shape[1]
        for k in (0, 4, 8):
            f[..., k::10] = w - e[..., k::10]
        f[..., 2::10], f[..., 3::10] = w - e[..., 6::10], e[..., 7::10]
        f[..., 6::10], f[..., 7::10] = w - e[..., 2::10], e[..., 3::10]
    elif direction == 'vertical':
        h = img_shape[0]
        for k in (3, 7, 9):
            f[..., k::10] = h - e[..., k::10]
        f[..., 0::10], f[..., 1::10] = e[..., 4::10], h - e[..., 5::10]
        f[..., 4::10], f[..., 5::10] = e[..., 0::10], h - e[..., 1::10]
    else:
        raise ValueError(f"Invalid flipping direction '{direction}'")
    return f


def keypoint_flip(k, img_shape, direction):
    """transforms.py:390-407: (G, 51) rows of [x, y, v]; horizontal mirrors x and swaps the left/right pairs."""
    f = k.copy()
    if direction == 'horizontal':
        f[:, 0::3] = img_shape[1] - f[:, 0::3]
        f = f.reshape(f.shape[0], f.shape[1] // 3, 3)       # (explicit count: an image without instances has 0 rows)
        for a, b in KEYPOINT_FLIP_PAIRS:
            f[:, [a, b]] = f[:, [b, a]]
        f = f.reshape(f.shape[0], k.shape[1])
    elif direction == 'vertical':
        f[:, 1::3] = img_shape[0] - f[:, 1::3]
    else:
        raise ValueError(f"Invalid flipping direction '{direction}'")
    return f


@PIPELINES.register_module()
class RandomFlip:
    """transforms.py:305-460."""

    def __init__(self, flip_ratio=None, direction='horizontal', keep_poly_clockwise=True):
        if flip_ratio is not None:
            assert 0 <= flip_ratio <= 1
        assert direction in ['horizontal', 'vertical']
        self.flip_ratio, self.direction, self.keep_poly_clockwise = flip_ratio, direction, keep_poly_clockwise

    def __call__(self, results):
        if 'flip' not in results:
            results['flip'] = bool(np.random.rand() < self.flip_ratio)
        if 'flip_direction' not in results:
            results['flip_direction'] = self.direction
        if results['flip']:
            d = results['flip_direction']
            for key in results.get('img_fields', ['img']):
                results[key] = np.flip(results[key], axis=1 if d == 'horizontal' else 0)
            for key in results.get('bbox_fields', []):
                results[key] = bbox_flip(results[key], results['img_shape'], d)
            for key in results.get('extreme_fields', []):
                results[key] = extreme_flip(results[key], results['img_shape'], d)
            for key in results.get('keypoint_fields', []):
                results[key] = keypoint_flip(results[key], results['img_shape'], d)
            for key in results.get('mask_fields', []):
                results[key] = results[key].flip(d, self.keep_poly_clockwise)
        return results


@PIPELINES.register_module()
class Normalize:
    """transforms.py:533-570 / mmcv/mmcv/image/photometric.py:5-41: float32 image, BGR -> RGB, (x − mean) · (1/std)
    with mean and 1/std held in float64."""

    def __init__(self, mean, std, to_rgb=True):
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)
        self.to_rgb = to_rgb

    def __call__(self, results):
        cv2 = _cv2()
        mean = np.float64(self.mean.reshape(1, -1))
        stdinv = 1 / np.float64(self.std.reshape(1, -1))
        for key in results.get('img_fields', ['img']):
            img = np.ascontiguousarray(results[key]).astype(np.float32)        # a copy: the passes below are in place
            if self.to_rgb:
                cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
            cv2.subtract(img, mean, img)         # OpenCV's SIMD passes, double scalars: the arithmetic the fixtures hold
            cv2.multiply(img, stdinv, img)
            results[key] = img
        results['img_norm_cfg'] = dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb)
        return results


@PIPELINES.register_module()
class Pad:
    """transforms.py:463-530: pad bottom / right to a fixed size or to the next multiple of ``size_divisor``."""

    def __init__(self, size=None, size_divisor=None, pad_val=0):
        assert (size is None) != (size_divisor is None)
        self.size, self.size_divisor, self.pad_val = size, size_divisor, pad_val

    def __call__(self, results):
        for key in results.get('img_fields', ['img']):
            img = results[key]
            if self.size is not None:
                ph, pw = self.size
            else:
                d = self.size_divisor
                ph, pw = -(-img.shape[0] // d) * d, -(-img.shape[1] // d) * d
            out = np.full((ph, pw) + img.shape[2:], self.pad_val, dtype=img.dtype)
            out[:img.shape[0], :img.shape[1]] = img
            results[key] = out
        results['pad_shape'] = out.shape
        results['pad_fixed_size'] = self.size
        results['pad_size_divisor'] = self.size_divisor
        for key in results.get('mask_fields', []):
            results[key] = results[key].pad(out.shape[:2], pad_val=self.pad_val)
        return results


@PIPELINES.register_module()
class DefaultFormatBundle:
    """formating.py:170-255 without the DataContainer wrapper (``collate`` below keys on the field name instead):
    image -> CHW tensor, ground-truth arrays -> tensors, polygon masks stay host objects."""

    TENSOR_KEYS = ('proposals', 'gt_bboxes', 'gt_bboxes_ignore', 'gt_labels', 'gt_extremes', 'gt_keypoints')

    def __call__(self, results):
        if 'img' in results:
            img = results['img']
            results.setdefault('pad_shape', img.shape)
            results.setdefault('scale_factor', 1.0)
            c = 1 if img.ndim < 3 else img.shape[2]
            results.setdefault('img_norm_cfg', dict(mean=np.zeros(c, np.float32), std=np.ones(c, np.float32),
                                                    to_rgb=False))
            if img.ndim < 3:
                img = img[..., None]
            results['img'] = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
        for key in self.TENSOR_KEYS:
            if key in results:
                results[key] = torch.from_numpy(np.ascontiguousarray(results[key]))
        return results


@PIPELINES.register_module()
class Collect:
    """formating.py:258-330: keep ``keys`` and gather ``meta_keys`` into ``img_metas``."""

    def __init__(self, keys, meta_keys=('filename', 'ori_filename', 'ori_shape', 'img_shape', 'pad_shape',
                                        'scale_factor', 'flip', 'flip_direction', 'img_norm_cfg')):
        self.keys, self.meta_keys = keys, meta_keys

    def __call__(self, results):
        data = {'img_metas': {k: results[k] for k in self.meta_keys if k in results}}
        for k in self.keys:
            data[k] = results[k]
        return data


@PIPELINES.register_module()
class ImageToTensor:
    """formating.py:66-99: HWC arrays under ``keys`` -> CHW tensors."""

    def __init__(self, keys):
        self.keys = keys

    def __call__(self, results):
        for key in self.keys:
            img = results[key]
            if img.ndim < 3:
                img = img[..., None]
            results[key] = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
        return results


@PIPELINES.register_module()
class MultiScaleFlipAug:
    """test_time_aug.py:9-117, the wrapper of every test pipeline in the LSNet configs: the inner transforms run once
    per (scale, flip, direction) with those three keys preset, and the per-augmentation dicts are transposed into one
    dict of lists (a single scale without flip gives lists of length 1 — ``LSDetector.forward_test`` takes them)."""

    def __init__(self, transforms, img_scale=None, scale_factor=None, flip=False, flip_direction='horizontal'):
        self.transforms = Compose(transforms)
        assert (img_scale is None) ^ (scale_factor is None), 'Must have but only one variable can be setted'
        if img_scale is not None:
            self.img_scale = img_scale if isinstance(img_scale, list) else [img_scale]
            self.scale_key = 'scale'
            assert all(isinstance(s, tuple) for s in self.img_scale), 'img_scale: a tuple or a list of tuples'
        else:
            self.img_scale = scale_factor if isinstance(scale_factor, list) else [scale_factor]
            self.scale_key = 'scale_factor'
        self.flip = flip
        self.flip_direction = flip_direction if isinstance(flip_direction, list) else [flip_direction]
        assert all(isinstance(d, str) for d in self.flip_direction), 'flip_direction: a str or a list of str'

    def __call__(self, results):
        aug = []
        for scale in self.img_scale:
            for flip in ([False, True] if self.flip else [False]):
                for direction in self.flip_direction:
                    r = results.copy()
                    r[self.scale_key] = scale
                    r['flip'] = flip
                    r['flip_direction'] = direction
                    aug.append(self.transforms(r))
        return {key: [d[key] for d in aug] for key in aug[0]}


@PIPELINES.register_module()
class DeviceFormatBundle:
    """The ``device_prep`` replacement for Normalize + Pad + DefaultFormatBundle: the resized (and flipped) uint8 HWC
    image goes out as it is; ``pad_shape`` / ``img_norm_cfg`` are recorded as the reference stages would have, and the
    arithmetic happens in ``lsnet_image_prep_u8`` on the GPU (``loader.DevicePrep``)."""

    def __init__(self, mean, std, to_rgb=True, size_divisor=32):
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)
        self.to_rgb, self.size_divisor = to_rgb, size_divisor

    def __call__(self, results):
        img = results['img']
        assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
        d = self.size_divisor
        results['pad_shape'] = (-(-img.shape[0] // d) * d, -(-img.shape[1] // d) * d, 3)
        results['pad_size_divisor'] = d
        results['img_norm_cfg'] = dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb)
        results['img'] = torch.from_numpy(np.ascontiguousarray(img))
        for key in DefaultFormatBundle.TENSOR_KEYS:
            if key in results:
                results[key] = torch.from_numpy(np.ascontiguousarray(results[key]))
        return results
