"""Batching for the train pipelines: aspect-ratio group samplers, collate, ``build_dataloader``.

Reference behaviour: mmdet/datasets/samplers/group_sampler.py:10-148 (every per-GPU batch comes from ONE aspect-ratio
group; the distributed sampler shuffles deterministically per epoch and pads every group to a multiple of
samples_per_gpu × world size), mmcv/mmcv/parallel/collate.py:11-85 (images are zero-padded bottom/right to the largest
of the batch and stacked; ground truth stays a per-image list), mmdet/datasets/builder.py:68-135 (``build_dataloader``,
worker seeding).  The collated batch is the dict ``LSDetector.forward_train`` / ``GraphTrainer.step`` take — no
DataContainer / scatter step in between, each process owns one GPU.
"""
import math
from functools import partial

import numpy as np
import torch
from torch.utils.data import DataLoader, Sampler

from ..registry import DATASETS, build_from_cfg


def build_dataset(cfg, default_args=None):
    """mmdet/datasets/builder.py:44-65 for the plain (un-wrapped) dataset types the LSNet configs use."""
    return build_from_cfg(cfg, DATASETS, default_args)


class GroupSampler(Sampler):
    """group_sampler.py:10-50 (single process): shuffle inside each group, pad the group to whole batches with random
    repeats, then shuffle the batches.  Draws from ``np.random`` in the reference's order."""

    def __init__(self, dataset, samples_per_gpu=1):
        assert hasattr(dataset, 'flag')
        self.samples_per_gpu = samples_per_gpu
        self.flag = dataset.flag.astype(np.int64)
        self.group_sizes = np.bincount(self.flag)
        self.num_samples = sum(int(np.ceil(s / samples_per_gpu)) * samples_per_gpu for s in self.group_sizes)

    def __iter__(self):
        spg = self.samples_per_gpu
        chunks = []
        for g, size in enumerate(self.group_sizes):
            if size == 0:
                continue
            idx = np.where(self.flag == g)[0]
            np.random.shuffle(idx)
            extra = int(np.ceil(size / spg)) * spg - len(idx)
            chunks.append(np.concatenate([idx, np.random.choice(idx, extra)]))
        flat = np.concatenate(chunks)
        order = np.random.permutation(range(len(flat) // spg))
        return iter(np.concatenate([flat[i * spg:(i + 1) * spg] for i in order]).astype(np.int64).tolist())

    def __len__(self):
        return self.num_samples


class DistributedGroupSampler(Sampler):
    """group_sampler.py:53-148: every rank builds the same epoch-seeded order and takes its contiguous slice, so the
    ranks' batches are disjoint and each batch stays inside one aspect-ratio group."""

    def __init__(self, dataset, samples_per_gpu=1, num_replicas=1, rank=0):
        assert hasattr(dataset, 'flag')
        self.samples_per_gpu, self.num_replicas, self.rank, self.epoch = samples_per_gpu, num_replicas, rank, 0
        self.flag = dataset.flag
        self.group_sizes = np.bincount(self.flag)
        self.num_samples = sum(int(math.ceil(s * 1.0 / samples_per_gpu / num_replicas)) * samples_per_gpu
                               for s in self.group_sizes)
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        g = torch.Generator()
        g.manual_seed(self.epoch)
        spg, n = self.samples_per_gpu, self.num_replicas
        indices = []
        for grp, size in enumerate(self.group_sizes):
            if size == 0:
                continue
            idx = np.where(self.flag == grp)[0]
            idx = idx[torch.randperm(int(size), generator=g).tolist()].tolist()
            extra = int(math.ceil(size * 1.0 / spg / n)) * spg * n - len(idx)
            indices.extend(idx + idx * (extra // size) + idx[:extra % size])
        assert len(indices) == self.total_size
        order = torch.randperm(len(indices) // spg, generator=g).tolist()
        indices = [indices[j] for i in order for j in range(i * spg, (i + 1) * spg)]
        lo = self.num_samples * self.rank
        return iter(indices[lo:lo + self.num_samples])

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


def collate(samples, pin=False, canvas_multiple=None):
    """collate.py:39-60 for one GPU's samples.  ``img`` CHW float tensors are zero-padded bottom/right to the batch
    maximum and stacked -> [B, 3, H, W]; uint8 HWC images of the device-prep pipeline are padded the same way ->
    [B, H, W, 3] with their valid extents in ``img_hw`` (the canvas is then the padded extents' maximum, exactly the
    stack of per-image ``Pad(size_divisor)`` results).  Everything else is a per-image list.
    ``canvas_multiple`` (not in the reference) rounds the batch canvas up to that multiple: with real COCO aspect
    ratios every multiple of 32 occurs as a canvas side, and ``GraphTrainer`` captures one CUDA graph per canvas —
    buckets of 128 keep that to a handful; each image's valid extent stays exact in ``img_metas['pad_shape']``, which
    is what the head's point validity uses (lsnet_head.py:770-779)."""
    def up(v):
        return v if not canvas_multiple else -(-v // canvas_multiple) * canvas_multiple
    if isinstance(samples[0]['img'], list):
        # test pipelines (MultiScaleFlipAug): every value is a list over augmentations -> one collated batch per
        # augmentation, returned as the dict of lists ``LSDetector.forward_test(imgs, img_metas)`` takes
        n_aug = len(samples[0]['img'])
        per_aug = [collate([{k: v[a] for k, v in s.items()} for s in samples], pin, canvas_multiple) for a in range(n_aug)]
        return {k: [b[k] for b in per_aug] for k in per_aug[0]}
    out = {k: [s[k] for s in samples] for k in samples[0] if k != 'img'}
    imgs = [s['img'] for s in samples]
    if imgs[0].dtype == torch.uint8:
        pads = [m['pad_shape'] for m in out['img_metas']]
        H, W = up(max(p[0] for p in pads)), up(max(p[1] for p in pads))
        batch = torch.zeros((len(imgs), H, W, 3), dtype=torch.uint8)
        for i, im in enumerate(imgs):
            batch[i, :im.shape[0], :im.shape[1]] = im
        out['img_hw'] = torch.tensor([[im.shape[0], im.shape[1]] for im in imgs], dtype=torch.int32)
        cfg = out['img_metas'][0]['img_norm_cfg']
        out['img_norm_cfg'] = cfg
    else:
        H, W = up(max(im.shape[-2] for im in imgs)), up(max(im.shape[-1] for im in imgs))
        batch = imgs[0].new_zeros((len(imgs), imgs[0].shape[0], H, W))
        for i, im in enumerate(imgs):
            batch[i, :, :im.shape[-2], :im.shape[-1]] = im
    out['img'] = batch.pin_memory() if pin else batch
    return out


def worker_init_fn(worker_id, num_workers, rank, seed):
    """builder.py:130-135 (seed of worker = num_workers * rank + worker_id + seed).  Additionally every worker runs
    OpenCV single-threaded: the parallelism of the loader is its worker processes, and a forked child must not touch
    the parent's OpenCV thread pool."""
    import random
    try:
        import cv2
        cv2.setNumThreads(0)
    except ImportError:
        pass
    if seed is not None:
        s = num_workers * rank + worker_id + seed
        np.random.seed(s)
        random.seed(s)


def device_prep_pipeline(pipeline):
    """Rewrite a reference train pipeline for GPU-side normalisation: ``Normalize`` and ``Pad(size_divisor)`` are
    dropped and ``DefaultFormatBundle`` becomes ``DeviceFormatBundle`` carrying their parameters.  Any other stage
    between them (none in the LSNet configs) would see different data, so it raises."""
    out, norm, pad = [], None, None
    for t in pipeline:
        kind = t['type']
        if kind == 'Normalize':
            norm = {k: v for k, v in t.items() if k != 'type'}
        elif kind == 'Pad':
            if t.get('size') is not None or t.get('pad_val', 0) != 0:
                raise ValueError('device_prep supports Pad(size_divisor=…, pad_val=0) only')
            pad = t.get('size_divisor')
        elif kind == 'DefaultFormatBundle':
            if norm is None or pad is None:
                raise ValueError('device_prep needs Normalize and Pad(size_divisor) before DefaultFormatBundle')
            out.append(dict(type='DeviceFormatBundle', size_divisor=pad, **norm))
        else:
            if norm is not None and kind != 'Collect':
                raise ValueError(f'device_prep: stage {kind} after Normalize would see un-normalised pixels')
            out.append(dict(t))
    return out


def build_dataloader(dataset, samples_per_gpu, workers_per_gpu, num_gpus=1, dist=True, shuffle=True, seed=None,
                     rank=None, world_size=None, pin=False, canvas_multiple=None, **kwargs):
    """builder.py:68-127.  One process per GPU: ``dist=True`` gives this rank's share through
    ``DistributedGroupSampler``; ``num_gpus`` other than 1 (the reference's single-process DataParallel mode) is not
    supported.  ``rank`` / ``world_size`` default to the initialised process group (mmcv ``get_dist_info``);
    ``canvas_multiple``: see ``collate``."""
    if rank is None or world_size is None:
        import torch.distributed as td
        on = td.is_available() and td.is_initialized()
        rank, world_size = (td.get_rank(), td.get_world_size()) if on else (0, 1)
    if num_gpus != 1:
        raise ValueError('one process per GPU: launch with torchrun instead of num_gpus > 1')
    if dist:
        if shuffle:
            sampler = DistributedGroupSampler(dataset, samples_per_gpu, world_size, rank)
        else:
            from torch.utils.data.distributed import DistributedSampler
            sampler = DistributedSampler(dataset, world_size, rank, shuffle=False)
    else:
        sampler = GroupSampler(dataset, samples_per_gpu) if shuffle else None
    init = partial(worker_init_fn, num_workers=workers_per_gpu, rank=rank, seed=seed)
    # pin=True: the DataLoader's pin thread of THIS process page-locks the collated tensors (the workers must not touch
    # CUDA), which is what lets GraphTrainer.prefetch copy the next batch asynchronously; the reference keeps
    # pin_memory=False and copies synchronously in scatter (mmcv/parallel/scatter_gather.py)
    return DataLoader(dataset, batch_size=samples_per_gpu, sampler=sampler, num_workers=workers_per_gpu,
                      collate_fn=partial(collate, canvas_multiple=canvas_multiple), pin_memory=bool(pin),
                      worker_init_fn=init, **kwargs)


class DevicePrep:
    """GPU half of the device-prep pipeline: uint8 [B, H, W, 3] batch (+ ``img_hw`` valid extents) -> normalised,
    zero-padded fp32 image in the layout the stem reads (NHWC memory = ``torch.channels_last`` [B, 3, H, W]) by ONE
    launch of ``lsnet_image_prep_u8``.  Matches Normalize -> Pad -> DefaultFormatBundle -> collate of the reference
    (transforms.py:463-570, formating.py:209-215, collate.py:39-60) bit for bit (double arithmetic, one rounding, as
    OpenCV's subtract / multiply with float64 scalars)."""

    def __init__(self, device='cuda'):
        self.device = torch.device(device)

    @staticmethod
    def constants(norm_cfg):
        mean = np.float64(np.asarray(norm_cfg['mean'], np.float32))       # photometric.py:35-36
        stdinv = 1 / np.float64(np.asarray(norm_cfg['std'], np.float32))
        return [float(v) for v in mean], [float(v) for v in stdinv]

    def run(self, u8, hw, norm_cfg, out=None):
        """``u8``: device uint8 [B, H, W, 3]; ``hw``: device int32 [B, 2]; ``out``: optional channels_last fp32
        [B, 3, H, W] to write into."""
        from ctypes import c_double as c_d
        from .. import lib as L
        B, H, W, _ = u8.shape
        if out is None:
            out = torch.empty((B, 3, H, W), device=u8.device, dtype=torch.float32).contiguous(
                memory_format=torch.channels_last)
        assert out.is_contiguous(memory_format=torch.channels_last) and out.dtype == torch.float32
        assert u8.is_contiguous() and u8.dtype == torch.uint8 and hw.dtype == torch.int32
        mean, stdinv = self.constants(norm_cfg)
        L.call('lsnet_image_prep_u8', L.ptr(u8), L.ptr(hw), L.c_int(B), L.c_int(H), L.c_int(W),
               c_d(mean[0]), c_d(mean[1]), c_d(mean[2]), c_d(stdinv[0]), c_d(stdinv[1]), c_d(stdinv[2]),
               L.c_int(1 if norm_cfg.get('to_rgb', True) else 0), L.ptr(out), L.stream())
        return out

    def __call__(self, batch):
        """Host batch of ``collate`` -> the same batch with ``img`` as the prepared device tensor."""
        u8 = batch['img'].to(self.device, non_blocking=True)
        hw = batch['img_hw'].to(self.device, non_blocking=True)
        out = dict(batch)
        out['img'] = self.run(u8, hw, batch['img_norm_cfg'])
        return out
