"""Train an LSNet config of the reference on COCO-format data with the B200 path: what the reference's
``tools/train.py`` + ``EpochBasedRunner`` do for these configs, reduced to the hot path (no hooks framework):

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_coco.py \\
        /path/to/configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py --work-dir work_dirs/bbox_r50 [--resume epoch_3.pth]

Config -> ``build_dataset(cfg.data.train)`` -> ``build_dataloader`` (aspect-ratio groups, one shard per rank; with
``--device-prep`` the workers ship uint8 images and the GPU normalises / pads) -> ``GraphTrainer`` (captured step, flat
buffers, one gradient all-reduce) -> ``train_epochs`` with one batch of look-ahead -> a checkpoint per epoch in the
reference's layout.  ``--dry-run N`` only walks N batches of the loader and prints what the step would receive (no GPU).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    ap.add_argument('config')
    ap.add_argument('--work-dir', default='work_dirs/lsnet')
    ap.add_argument('--epochs', type=int, default=None, help='default: cfg.total_epochs')
    ap.add_argument('--resume', default=None)
    ap.add_argument('--device-prep', action='store_true', help='uint8 batches, Normalize + Pad on the GPU')
    ap.add_argument('--canvas-multiple', type=int, default=128, help='canvas buckets of the CUDA-graph cache')
    ap.add_argument('--ann-file', default=None, help='override cfg.data.train.ann_file')
    ap.add_argument('--img-prefix', default=None, help='override cfg.data.train.img_prefix')
    ap.add_argument('--samples-per-gpu', type=int, default=None)
    ap.add_argument('--workers-per-gpu', type=int, default=None)
    ap.add_argument('--log-interval', type=int, default=50)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--dry-run', type=int, default=0, metavar='N')
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist
    from lsnet_b200 import Config, datasets as D
    cfg = Config.fromfile(args.config)
    train = dict(cfg.data.train)
    if args.ann_file:
        train['ann_file'] = args.ann_file
    if args.img_prefix is not None:
        train['img_prefix'] = args.img_prefix
    if args.device_prep:
        train['pipeline'] = D.device_prep_pipeline(train['pipeline'])
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not args.dry_run:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    ds = D.build_dataset(train)
    dl = D.build_dataloader(ds, args.samples_per_gpu or cfg.data.samples_per_gpu,
                            cfg.data.workers_per_gpu if args.workers_per_gpu is None else args.workers_per_gpu,
                            dist=True, seed=args.seed, rank=rank, world_size=world, pin=not args.dry_run,
                            canvas_multiple=args.canvas_multiple)
    if rank == 0:
        print(f'{type(ds).__name__}: {len(ds)} images, {len(dl)} iterations per epoch and rank, '
              f'{len(ds.CLASSES)} classes, pipeline {[type(t).__name__ for t in ds.pipeline.transforms]}')
    if args.dry_run:
        for i, b in enumerate(dl):
            if i >= args.dry_run:
                break
            gt = [k for k in b if k.startswith('gt_')]
            print(f'batch {i}: img {tuple(b["img"].shape)} {b["img"].dtype}, '
                  f'instances {[len(x) for x in b["gt_bboxes"]]}, ground truth {gt}')
        return 0

    from lsnet_b200.train import GraphTrainer, resume, save_checkpoint, train_epochs
    epochs = args.epochs or cfg.total_epochs
    torch.manual_seed(args.seed)
    tr = GraphTrainer(cfg, next(iter(dl)), device=f'cuda:{torch.cuda.current_device()}', distributed=world > 1,
                      iters_per_epoch=len(dl))
    start = 0
    if args.resume:
        start = int(resume(tr, args.resume).get('epoch', 0))
    os.makedirs(args.work_dir, exist_ok=True)
    t0 = [time.perf_counter()]

    def log(epoch, i, loss, log_vars):
        if rank == 0 and (i + 1) % args.log_interval == 0:
            dt, t0[0] = time.perf_counter() - t0[0], time.perf_counter()
            terms = ', '.join(f'{k}: {float(v):.4f}' for k, v in log_vars.items())      # one device sync per interval
            print(f'Epoch [{epoch + 1}][{i + 1}/{len(dl)}] lr: {tr.lr_at(tr.iter):.5f}, '
                  f'time: {dt / args.log_interval:.3f}, {terms}', flush=True)
    for epoch in range(start, epochs):
        train_epochs(tr, dl, epochs=1, start_epoch=epoch, on_step=log)
        if rank == 0:
            save_checkpoint(tr, os.path.join(args.work_dir, f'epoch_{epoch + 1}.pth'), epoch=epoch + 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
