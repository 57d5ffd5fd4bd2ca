"""Host-side cost of the input pipeline per image (one core): the UNMODIFIED reference pipeline (imported from
/root/reference through oracle/ref_harness.py -- this container only) beside lsnet_b200.datasets, for the three tasks at
the real training scale (COCO-sized 480x640 JPEGs -> Resize (1333, 800) -> flip -> Normalize -> Pad -> bundle), and the
worker half of the device-prep variant (Normalize / Pad / transpose move to lsnet_image_prep_u8 on the GPU).
Writes profiles/r02_datapath_cpu.md.   usage: python tools/bench_datapath.py [n_images]"""
import copy
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cv2  # noqa: E402

cv2.setNumThreads(1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)


def make_coco(tmp, pose):
    rng = np.random.RandomState(3)
    images, anns, aid = [], [], 1
    for i in range(N):
        h, w = (480, 640) if i % 4 else (640, 480)
        img = cv2.GaussianBlur(rng.randint(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3)
        cv2.imwrite(os.path.join(tmp, f'{i}.jpg'), img, [cv2.IMWRITE_JPEG_QUALITY, 90])
        images.append(dict(id=i, file_name=f'{i}.jpg', height=h, width=w))
        for j in range(7):                                   # COCO: 7.3 instances per image on average
            bw, bh = rng.uniform(20, w * 0.6), rng.uniform(20, h * 0.6)
            x1, y1 = rng.uniform(0, w - bw), rng.uniform(0, h - bh)
            n = int(rng.randint(12, 90))
            th = np.sort(rng.rand(n)) * 2 * np.pi
            r = 0.6 + 0.4 * rng.rand(n)
            seg = np.stack([x1 + bw / 2 + bw / 2 * r * np.cos(th), y1 + bh / 2 + bh / 2 * r * np.sin(th)], 1).reshape(-1)
            a = dict(id=aid, image_id=i, category_id=1, bbox=[x1, y1, bw, bh], area=bw * bh * 0.6, iscrowd=0,
                     segmentation=[seg.round(2).tolist()])
            if pose:
                k = np.stack([x1 + rng.rand(17) * bw, y1 + rng.rand(17) * bh, rng.randint(1, 3, 17)], 1)
                a['keypoints'] = k.reshape(-1).round(2).tolist()
            else:
                a['extreme_points'] = [x1 + bw / 2, y1, x1, y1 + bh / 2, x1 + bw / 2, y1 + bh, x1 + bw, y1 + bh / 2,
                                       x1 + bw / 2, y1 + bh / 2]
            anns.append(a)
            aid += 1
    return dict(images=images, annotations=anns, categories=[dict(id=1, name='person')])


def pipeline(task):
    import synth_coco as S
    p = S.pipeline(task)
    for t in p:
        if t['type'] == 'Resize':
            t['img_scale'] = (1333, 800)
    return [dict(type='LoadImageFromFile')] + p


def timed(fn, items):
    fn(items[0])
    t0 = time.perf_counter()
    for it in items:
        fn(it)
    return 1e3 * (time.perf_counter() - t0) / len(items)


def main():
    from oracle import ref_harness as rh
    rh.load()
    import make_golden_data as M
    import mmdet.datasets.pipelines.loading as loading
    from mmdet.datasets.pipelines import Compose as RefCompose
    loading.Polygon = M._Polygon
    from lsnet_b200 import datasets as D
    from lsnet_b200.registry import DATASETS
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        for task in ('bbox', 'segm', 'pose_bbox'):
            pose = task == 'pose_bbox'
            coco = make_coco(tmp, pose)
            ds = DATASETS.get('CocoPoseDataset' if pose else 'CocoDataset')(ann_file=coco, pipeline=pipeline(task),
                                                                            img_prefix=tmp)
            dev = DATASETS.get('CocoPoseDataset' if pose else 'CocoDataset')(
                ann_file=coco, pipeline=D.device_prep_pipeline(pipeline(task)), img_prefix=tmp)
            ref = RefCompose(pipeline(task))

            def base(i):
                r = dict(img_info=ds.data_infos[i], ann_info=copy.deepcopy(ds.get_ann_info(i)))
                ds.pre_pipeline(r)
                return r
            idx = list(range(len(ds)))
            np.random.seed(0)
            t_ref = timed(lambda i: ref(base(i)), idx)
            np.random.seed(0)
            t_own = timed(lambda i: ds.pipeline(base(i)), idx)
            np.random.seed(0)
            t_dev = timed(lambda i: dev.pipeline(base(i)), idx)
            rows.append((task, t_ref, t_own, t_dev))
            print(task, f'reference {t_ref:.1f} ms, lsnet_b200 {t_own:.1f} ms, device-prep workers {t_dev:.1f} ms per image')
    out = os.path.join(ROOT, 'profiles', 'r02_datapath_cpu.md')
    with open(out, 'w') as f:
        f.write('# Input pipeline on the host: ms per image on ONE core of the build container\n\n'
                f'`python tools/bench_datapath.py {N}`: {N} COCO-sized JPEGs (480x640 / 640x480, quality 90), 7 instances each '
                '(polygons of 12-90 vertices), LoadImageFromFile -> LoadAnnotations -> Resize (1333, 800) -> RandomFlip(0.5) '
                '-> Normalize -> Pad(32) -> DefaultFormatBundle -> Collect; OpenCV single-threaded, the same seeds for every '
                'column.  Outputs are bit-identical across the first two columns (tests/test_datasets_host.py); the third '
                'stops at the resized uint8 image and leaves Normalize + Pad + transpose to `lsnet_image_prep_u8` on the GPU '
                '(17.8 us per 4-image batch, `r02_ncu_image_prep_u8.md`).\n\n'
                '| task | reference pipeline | lsnet_b200.datasets | device-prep workers | images/s/core (reference -> device-prep) |\n'
                '|---|---:|---:|---:|---:|\n')
        for task, a, b, c in rows:
            f.write(f'| {task} | {a:.1f} ms | {b:.1f} ms | {c:.1f} ms | {1e3 / a:.0f} -> {1e3 / c:.0f} |\n')
        f.write('\nAt 180 images/s per GPU (`r02_bench_bbox_r50.json`) one B200 consumes the output of '
                + ', '.join(f'{180 * c / 1e3:.1f} ({task})' for task, _, _, c in rows)
                + ' loader cores with the device-prep pipeline against '
                + ', '.join(f'{180 * a / 1e3:.1f}' for _, a, _, _ in rows) + ' with the reference pipeline.\n')
    print(open(out).read())


if __name__ == '__main__':
    main()
