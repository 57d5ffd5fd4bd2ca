"""Generates tests/golden/datapath.npz: outputs of the UNMODIFIED reference data path (mmdet.datasets pipelines, COCO
annotation parsers, group samplers, mmcv collate — imported from /root/reference through oracle/ref_harness.py) on the
seeded annotations / images of synth_coco.py.  Only reference OUTPUTS are stored; the tests rebuild the inputs.

    python tests/golden/make_golden_data.py        # needs /root/reference (this container only)

Two stand-ins, both outside the code under test: shapely is not installed, so ``Polygon(p).exterior.is_ccw`` is served
by the ring's signed area (shapely's own definition); pycocotools is not installed, so the parsers are called as
``CocoDataset._parse_ann_info(stub_self, img_info, anns)`` with the two attributes they read (cat_ids, cat2label).
"""
import copy
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth_coco as S  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'datapath.npz')


class _Ring:
    def __init__(self, p):
        x, y = p[:, 0], p[:, 1]
        self.is_ccw = bool(0.5 * (np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))) > 0)


class _Polygon:
    def __init__(self, p):
        self.exterior = _Ring(np.asarray(p, dtype=np.float64))


def main():
    rh.load()
    import mmdet.datasets.pipelines.loading as loading
    from mmcv.parallel import collate
    from mmdet.datasets import CocoDataset
    from mmdet.datasets.coco_pose import CocoPoseDataset
    from mmdet.datasets.pipelines import Compose
    from mmdet.datasets.samplers import DistributedGroupSampler, GroupSampler
    loading.Polygon = _Polygon
    rec = {}

    # ---- annotation parsers ----
    parsed = {}
    for pose in (False, True):
        d = S.coco_dict(pose)
        cls = CocoPoseDataset if pose else CocoDataset
        names = cls.CLASSES
        cat_ids = [c['id'] for c in d['categories'] if c['name'] in names]
        stub = types.SimpleNamespace(cat_ids=cat_ids, cat2label={c: i for i, c in enumerate(cat_ids)})
        by_img = {}
        for a in d['annotations']:
            by_img.setdefault(a['image_id'], []).append(a)
        for i, info in enumerate(d['images'][:len(S.SIZES)]):
            info = dict(info, filename=info['file_name'])
            ann = cls._parse_ann_info(stub, info, by_img.get(info['id'], []))
            parsed[(pose, i)] = (info, ann)
            tag = f'parse_{"pose" if pose else "det"}_{i}'
            rec[tag + '_bboxes'] = ann['bboxes']
            rec[tag + '_labels'] = ann['labels']
            rec[tag + '_ignore'] = ann['bboxes_ignore']
            rec[tag + ('_keypoints' if pose else '_extremes')] = ann['keypoints' if pose else 'extremes']
            rec[tag + '_nmask'] = np.array(len(ann['masks']))

    # ---- pipelines ----
    samples = {}
    for task in ('bbox', 'segm', 'pose_bbox'):
        for ms in (0, 1):
            pipe = Compose(S.pipeline(task, bool(ms)))
            for i in range(len(S.SIZES)):
                info, ann = parsed[(task == 'pose_bbox', i)]
                img = S.image(i)
                res = dict(img_info=info, ann_info=copy.deepcopy(ann), img=img, img_shape=img.shape, ori_shape=img.shape,
                           img_fields=['img'], filename=info['filename'], ori_filename=info['filename'], img_prefix=None,
                           bbox_fields=[], extreme_fields=[], mask_fields=[], seg_fields=[], keypoint_fields=[])
                np.random.seed(1000 + 10 * i + ms)
                out = pipe(res)
                samples[(task, ms, i)] = out
                tag = f'pipe_{task}_{ms}_{i}'
                im = out['img'].data.numpy()
                if task == 'bbox':
                    rec[tag + '_img'] = im
                rec[tag + '_imgsum'] = np.array([im.astype(np.float64).sum(), np.abs(im.astype(np.float64)).sum()])
                meta = out['img_metas'].data
                rec[tag + '_meta'] = np.array(list(meta['img_shape']) + list(meta['pad_shape']) + [int(meta['flip'])],
                                              np.int64)
                rec[tag + '_scale_factor'] = np.asarray(meta['scale_factor'], np.float32)
                rec[tag + '_gt_bboxes'] = out['gt_bboxes'].data.numpy()
                rec[tag + '_gt_labels'] = out['gt_labels'].data.numpy()
                if task == 'bbox':
                    rec[tag + '_gt_extremes'] = out['gt_extremes'].data.numpy()
                elif task == 'pose_bbox':
                    rec[tag + '_gt_keypoints'] = out['gt_keypoints'].data.numpy()
                else:
                    pm = out['gt_masks'].data
                    rec[tag + '_mask_hw'] = np.array([pm.height, pm.width])
                    rec[tag + '_mask_ncomp'] = np.array([len(c) for c in pm.masks])
                    rec[tag + '_mask_pts'] = np.concatenate([p for c in pm.masks for p in c]) if len(pm.masks) else \
                        np.zeros(0)

    # ---- test pipelines (MultiScaleFlipAug) ----
    for multi in (0, 1):
        pipe = Compose(S.eval_pipeline(bool(multi)))
        for i in (0, 1):
            info, _ = parsed[(False, i)]
            img = S.image(i)
            out = pipe(dict(img_info=info, img=img, img_shape=img.shape, ori_shape=img.shape, img_fields=['img'],
                            filename=info['filename'], ori_filename=info['filename'], img_prefix=None, bbox_fields=[],
                            extreme_fields=[], mask_fields=[], seg_fields=[], keypoint_fields=[]))
            tag = f'test_{multi}_{i}'
            rec[tag + '_naug'] = np.array(len(out['img']))
            for a, (im, meta) in enumerate(zip(out['img'], out['img_metas'])):
                m = meta.data
                rec[f'{tag}_{a}_img'] = im.numpy() if (multi == 0 or a == 3) else np.zeros(0, np.float32)
                rec[f'{tag}_{a}_imgsum'] = np.array([im.double().sum().item(), im.double().abs().sum().item()])
                rec[f'{tag}_{a}_meta'] = np.array(list(m['img_shape']) + list(m['pad_shape']) + [int(m['flip'])], np.int64)
                rec[f'{tag}_{a}_scale_factor'] = np.asarray(m['scale_factor'], np.float32)

    # ---- multi-scale testing by voting: the host half of aug_test_vote (lsnet.py:138-365) on synthetic detections ----
    import torch
    from mmdet.models.detectors.lsnet import LSDetector
    torch.Tensor.cuda = lambda self, *a, **k: self                 # instances_vote ends in .cuda(): stay on the CPU
    for task in ('bbox', 'segm', 'pose_bbox'):
        dets, metas = S.tta_case(task)
        stub = types.SimpleNamespace(bbox_head=types.SimpleNamespace(task=task, num_classes=3))
        ab, av, al = [], [], []
        for i, (b, v, l) in enumerate(dets):                       # lsnet.py:313-320
            b, v, l = torch.from_numpy(b), torch.from_numpy(v), torch.from_numpy(l)
            keep = LSDetector.remove_boxes(stub, b, *S.TTA_SCALE_RANGES[i // 2])
            rec[f'tta_{task}_keep_{i}'] = keep.numpy()
            ab.append(b[keep, :]); av.append(v[keep, :]); al.append(l[keep])
        mb, mv, ml = LSDetector.merge_aug_vote_results(stub, ab, av, al, [[m] for m in metas])      # :323-324
        rec[f'tta_{task}_mapped_boxes'], rec[f'tta_{task}_mapped_vectors'] = mb.numpy(), mv.numpy()
        ob, ov, ol = [], [], []
        for j in range(3):                                         # :329-343
            inds = (ml == j).nonzero().squeeze(1)
            bj, vj, sj = LSDetector.instances_vote(stub, mb[inds, :4].view(-1, 4), mv[inds], mb[inds, 4])
            if len(bj) > 0:
                ob.append(torch.cat([bj, sj[:, None]], dim=1)); ov.append(vj)
                ol.append(torch.full((bj.shape[0],), j, dtype=torch.int64))
        rec[f'tta_{task}_boxes'] = torch.cat(ob).numpy()
        rec[f'tta_{task}_vectors'] = torch.cat(ov).numpy()
        rec[f'tta_{task}_labels'] = torch.cat(ol).numpy()

    # ---- the head's host-side polygon step on pipeline output (lsnet_head.py:1717-1756) ----
    from mmdet.models.dense_heads.lsnet_head import LSHead
    stub = types.SimpleNamespace(component_polygon_area=lambda poly: LSHead.component_polygon_area(None, poly))
    masks = [samples[('segm', 1, i)]['gt_masks'].data for i in range(len(S.SIZES))]
    polys, boxes = LSHead.process_polygons(stub, masks, [torch.zeros(1)])
    for i, (p, b) in enumerate(zip(polys, boxes)):
        rec[f'headpoly_{i}_table'] = p.numpy()
        rec[f'headpoly_{i}_boxes'] = b.numpy()

    # ---- collate (mmcv DataContainer semantics) of mixed-size samples ----
    for name, ids in (('a', [0, 3]), ('b', [1, 4, 2])):
        b = collate([samples[('bbox', 1, i)] for i in ids], samples_per_gpu=len(ids))
        rec[f'collate_{name}_shape'] = np.array(b['img'].data[0].shape)
        rec[f'collate_{name}_sum'] = np.array(b['img'].data[0].double().sum().item())

    # ---- samplers ----
    flag = np.array([1, 0, 1, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0], np.uint8)
    ds = types.SimpleNamespace(flag=flag)
    np.random.seed(5)
    rec['group_sampler'] = np.array(list(GroupSampler(ds, samples_per_gpu=2)))
    for epoch in (0, 4):
        for rank in range(3):
            s = DistributedGroupSampler(ds, samples_per_gpu=2, num_replicas=3, rank=rank)
            s.set_epoch(epoch)
            rec[f'dist_sampler_e{epoch}_r{rank}'] = np.array(list(s))

    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KB,', len(rec), 'arrays')


if __name__ == '__main__':
    main()
