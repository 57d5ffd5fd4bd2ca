"""Run an LSNet config of the reference over a COCO-format test set with the B200 path: what the reference's
``tools/test.py`` + ``single_gpu_test`` do (mmdet/apis/test.py:12-60), reduced to the path:

    python tools/test_coco.py /path/to/configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py work_dirs/bbox_r50/epoch_12.pth \\
        --out results.pkl [--vote '[[0, 10000]]' --scales '[(1333, 800)]' --flip]

Config -> ``build_dataset(cfg.data.test, test_mode=True)`` -> the config's own MultiScaleFlipAug test pipeline -> one image
per step through ``LSDetector.forward_test`` (device decode + NMS; several augmentations: instance voting) with
``rescale=True`` -> the reference's per-image result lists ``[boxes per class, landmark vectors per class]``, pickled.
``--dry-run N`` only walks N samples of the loader and prints what the model would receive (no GPU, no checkpoint).
"""
import argparse
import ast
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    ap.add_argument('config')
    ap.add_argument('checkpoint', nargs='?')
    ap.add_argument('--out', default='results.pkl')
    ap.add_argument('--ann-file', default=None)
    ap.add_argument('--img-prefix', default=None)
    ap.add_argument('--scales', default=None, help="python list of (long, short) scales replacing the pipeline's img_scale")
    ap.add_argument('--flip', action='store_true', help='also test the flipped image at every scale')
    ap.add_argument('--vote', default=None, help="python list of [min, max] box-size ranges, one per scale: "
                                                 "test_cfg.method='vote' (multi-scale testing of the reference's tables)")
    ap.add_argument('--workers', type=int, default=2)
    ap.add_argument('--dry-run', type=int, default=0, metavar='N')
    args = ap.parse_args(argv)

    import torch
    from lsnet_b200 import Config, datasets as D
    cfg = Config.fromfile(args.config)
    test = dict(cfg.data.test)
    if args.ann_file:
        test['ann_file'] = args.ann_file
    if args.img_prefix is not None:
        test['img_prefix'] = args.img_prefix
    pipe = [dict(t) for t in test['pipeline']]
    for t in pipe:
        if t['type'] == 'MultiScaleFlipAug':
            if args.scales:
                t['img_scale'] = [tuple(s) for s in ast.literal_eval(args.scales)]
            if args.flip:
                t['flip'] = True
    test['pipeline'] = pipe
    test['test_mode'] = True
    ds = D.build_dataset(test)
    dl = D.build_dataloader(ds, 1, args.workers, dist=False, shuffle=False)
    print(f'{type(ds).__name__}: {len(ds)} images, test pipeline {[type(t).__name__ for t in ds.pipeline.transforms]}')
    if args.dry_run:
        for i, b in enumerate(dl):
            if i >= args.dry_run:
                break
            print(f'image {i}: {len(b["img"])} augmentation(s), sizes {[tuple(x.shape[-2:]) for x in b["img"]]}, '
                  f'flips {[m[0]["flip"] for m in b["img_metas"]]}')
        return 0

    from lsnet_b200 import build_detector
    test_cfg = dict(cfg.test_cfg)
    if args.vote:
        test_cfg.update(method='vote', scale_ranges=ast.literal_eval(args.vote))
    cfg.model.pretrained = None
    model = build_detector(cfg.model, train_cfg=None, test_cfg=test_cfg)
    from lsnet_b200.train import load_checkpoint_file
    ck = load_checkpoint_file(args.checkpoint)
    state = ck.get('state_dict', ck)
    model.load_state_dict({(k[7:] if k.startswith('module.') else k): v for k, v in state.items()}, strict=True)
    model.cuda().eval()
    results = []
    for b in dl:
        with torch.no_grad():
            r = model(img=[x.cuda(non_blocking=True) for x in b['img']], img_metas=b['img_metas'], return_loss=False,
                      rescale=True)
        results.append(r[0] if len(b['img']) == 1 else r)       # simple_test returns a list over the batch
    with open(args.out, 'wb') as f:
        pickle.dump(results, f)
    files, _ = ds.format_results(results, jsonfile_prefix=os.path.splitext(args.out)[0])      # COCO json for the COCO API
    print(f'wrote {len(results)} results to {args.out} and {sorted(set(files.values()))}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
