"""The training step of the LSNet hot path: forward + loss + backward + grad-clip + SGD(momentum), data-parallel
through torch DDP over NCCL (one process per GPU).  Mirrors what ``OptimizerHook.after_train_iter`` +
``EpochBasedRunner.train`` do per iteration (mmcv/mmcv/runner/hooks/optimizer.py:19-28,
mmcv/mmcv/runner/epoch_based_runner.py:20-47) without the reference's per-scalar all-reduce + .item() every step."""
import os

import torch

from . import lib as L
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

from .modules.detector import parse_losses
from .registry import build_detector


def warmup_lr(base_lr, it, warmup_iters=500, warmup_ratio=0.001):
    """Linear warm-up of configs/_base_/schedules/schedule_1x.py:5-11 (mmcv lr_updater.py:123-141)."""
    if it >= warmup_iters:
        return base_lr
    k = (1 - it / warmup_iters) * (1 - warmup_ratio)
    return base_lr * (1 - k)


def grad_clip_of(cfg):
    """``max_norm`` of the gradient clip: the reference configs carry it as
    ``optimizer_config = dict(grad_clip=dict(max_norm=35, norm_type=2))`` (configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py:65,
    consumed by OptimizerHook, mmcv/runner/hooks/optimizer.py:10-17); a top-level ``grad_clip`` is accepted too."""
    clip = (cfg.get('optimizer_config') or {}).get('grad_clip') or cfg.get('grad_clip') or {}
    nt = clip.get('norm_type', 2)
    if clip and nt != 2:
        raise ValueError(f'grad_clip norm_type={nt}: only the L2 norm is implemented')
    return clip.get('max_norm')


class LrSchedule:
    """``lr_config = dict(policy='step', warmup='linear', warmup_iters=500, warmup_ratio=0.001, step=[8, 11])`` of
    configs/_base_/schedules/schedule_1x.py (mmcv StepLrUpdaterHook + LrUpdaterHook.get_warmup_lr,
    mmcv/runner/hooks/lr_updater.py:62-141, 144-172): the regular lr decays by ``gamma`` at each epoch in ``step``; during
    the first ``warmup_iters`` iterations the (already decayed) regular lr is scaled linearly from ``warmup_ratio``.
    Epochs need ``iters_per_epoch``; without it the lr never decays."""

    def __init__(self, base_lr, lr_config=None, iters_per_epoch=None):
        c = dict(lr_config or {})
        policy = c.get('policy', 'step')
        if policy not in ('step', 'fixed'):
            raise ValueError(f'lr_config policy {policy!r} is not implemented (step / fixed)')
        self.base_lr = base_lr
        self.steps = sorted([c['step']] if isinstance(c.get('step'), int) else list(c.get('step') or [])) if policy == 'step' else []
        self.gamma = c.get('gamma', 0.1)
        self.warmup = c.get('warmup', 'linear')
        if self.warmup not in (None, 'linear', 'constant', 'exp'):
            raise ValueError(f'warmup {self.warmup!r}')
        self.warmup_iters, self.warmup_ratio = c.get('warmup_iters', 500), c.get('warmup_ratio', 0.001)
        self.iters_per_epoch = iters_per_epoch

    def __call__(self, it):
        lr = self.base_lr
        if self.iters_per_epoch and self.steps:
            epoch = it // self.iters_per_epoch
            lr = lr * self.gamma ** sum(1 for s in self.steps if epoch >= s)
        if self.warmup is None or it >= self.warmup_iters:
            return lr
        if self.warmup == 'constant':
            return lr * self.warmup_ratio
        if self.warmup == 'exp':
            return lr * self.warmup_ratio ** (1 - it / self.warmup_iters)
        return lr * (1 - (1 - it / self.warmup_iters) * (1 - self.warmup_ratio))


class Trainer:

    def __init__(self, cfg, device='cuda', distributed=False, bucket_cap_mb=64, model=None, iters_per_epoch=None):
        self.device = torch.device(device)
        self.model = model if model is not None else build_detector(cfg['model'], train_cfg=cfg.get('train_cfg'),
                                                                    test_cfg=cfg.get('test_cfg'))
        self.model.to(self.device).train()
        self.core = self.model
        if distributed:
            ids = [self.device.index] if self.device.type == 'cuda' else None
            self.model = DDP(self.core, device_ids=ids, broadcast_buffers=False, bucket_cap_mb=bucket_cap_mb,
                             gradient_as_bucket_view=True)
        opt = cfg.get('optimizer', dict(lr=0.01, momentum=0.9, weight_decay=1e-4))
        self.base_lr = opt['lr']
        params = [p for p in self.core.parameters() if p.requires_grad]
        self.params = params
        self.optimizer = torch.optim.SGD(params, lr=opt['lr'], momentum=opt.get('momentum', 0.9),
                                         weight_decay=opt.get('weight_decay', 1e-4), foreach=True)
        self.max_norm = grad_clip_of(cfg)
        self.lr_at = LrSchedule(self.base_lr, cfg.get('lr_config'), iters_per_epoch)
        self.iter = 0

    def step(self, batch, sync_log=False):
        """One training iteration on a per-GPU batch whose image is already on the device.  Returns the loss tensor
        (device) and log_vars (device tensors, or python floats when sync_log)."""
        for g in self.optimizer.param_groups:
            g['lr'] = self.lr_at(self.iter)
        self.optimizer.zero_grad(set_to_none=True)
        losses = self.model(**batch)
        loss, log_vars = parse_losses(losses, sync_log)
        loss.backward()
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_norm, foreach=True)
        self.optimizer.step()
        self.iter += 1
        return loss, log_vars


DIRECT_WGRAD = os.environ.get('LSNET_DIRECT_WGRAD', '1') == '1'
DIRECT_VEC = os.environ.get('LSNET_DIRECT_VEC', '1') == '1'


class GraphTrainer:
    """The same training step with (almost) no host work per iteration, B200-first:

      * forward + loss + backward are captured ONCE into a CUDA graph over static input buffers (image, PackedGT);
        each step = refresh the static buffers (async copies) + one graph replay;
      * every trainable parameter / gradient / momentum lives in ONE flat fp32 buffer, so data parallelism is a single
        NCCL all-reduce of the flat gradient (154 MB; ~0.4 ms over NVLink 5 / NVSwitch — no bucketing or overlap
        machinery needed at this bandwidth) and clip + SGD(momentum, wd) are a handful of flat element-wise kernels.

    Semantics are those of ``Trainer`` / the reference's OptimizerHook: grads averaged over ranks, L2 clip at
    ``max_norm`` on the averaged gradient, torch.optim.SGD update (dampening 0, no nesterov)."""

    def __init__(self, cfg, sample_batch, device='cuda', distributed=False, capacity=128, model=None,
                 kernel_timing=False, iters_per_epoch=None, max_graphs=8):
        """``capacity``: ground-truth instances per image the static buffers hold (COCO images carry up to ~100); a
        batch with more makes ``load_batch`` grow the buffers and re-capture the graph."""
        self.device = torch.device(device)
        self.distributed = distributed
        self.model = model if model is not None else build_detector(cfg['model'], train_cfg=cfg.get('train_cfg'),
                                                                    test_cfg=cfg.get('test_cfg'))
        self.model.to(self.device).train()
        self.core = self.model
        opt = cfg.get('optimizer', dict(lr=0.01, momentum=0.9, weight_decay=1e-4))
        self.base_lr, self.momentum, self.wd = opt['lr'], opt.get('momentum', 0.9), opt.get('weight_decay', 1e-4)
        self.max_norm = grad_clip_of(cfg)
        self.lr_at = LrSchedule(self.base_lr, cfg.get('lr_config'), iters_per_epoch)
        self.capacity = capacity
        self.kernel_timing = kernel_timing      # capture external event-record nodes around the library's kernels
        self.iter = 0
        if distributed:     # identical replicas: broadcast rank 0's initial parameters/buffers once
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, 0)
        # ---- flat parameter / gradient / momentum storage ----
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        al = 64                                  # every parameter starts on a 256-byte boundary of the flat buffers
        n = sum((p.numel() + al - 1) // al * al for p in self.params)
        self.flat_p = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=self.device, dtype=torch.float32)
        o = 0
        self._slots = []                         # (offset, tap-major?) of every parameter inside the flat buffers
        for p in self.params:
            k = p.numel()
            self._slots.append((o, bool(DIRECT_WGRAD and p.dim() == 4 and getattr(p, '_lsnet_tapmajor', False))))
            if DIRECT_WGRAD and p.dim() == 4 and getattr(p, '_lsnet_tapmajor', False):
                # Weights of the tcgen05 conv / DCN kernels live tap-major ([Cout][kh][kw][Cin] memory, i.e.
                # torch.channels_last strides on the OIHW parameter): the bf16 operand pack is then a plain cast and the
                # weight-gradient GEMM's [Cout, taps*Cin] fp32 output IS the gradient's memory, so it accumulates
                # there directly (no memset, no permute, no AccumulateGrad add per pyramid level).
                co, ci, kh, kw = p.shape
                self.flat_p[o:o + k].view(co, kh, kw, ci).copy_(p.data.permute(0, 2, 3, 1))
                p.data = self.flat_p[o:o + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
                p.grad = self.flat_g[o:o + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
                p._lsnet_grad2d = self.flat_g[o:o + k].view(co, kh * kw * ci)
            else:
                self.flat_p[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[o:o + k].view_as(p)
                p.grad = self.flat_g[o:o + k].view_as(p)
                # biases / norm affines: the backward kernels add into this memory directly (ops.gemm_ops.direct_vec)
                p._lsnet_direct_vec = DIRECT_VEC and p.dim() == 1
            p._lsnet_direct_any = DIRECT_VEC
            o += (k + al - 1) // al * al
        # ---- captured steps, one per input canvas (H, W): multi-scale training (configs/lsnet/*mstrain*) replays the
        # graph of the batch's shape bucket; all graphs share ONE memory pool (they never run concurrently), so the
        # footprint is that of the largest shape ----
        self.steps = {}
        self.max_graphs = max_graphs
        self._pool = None
        self.recaptures = 0
        self.cur = self._entry(sample_batch)

    def flat_views(self, flat):
        """Per-parameter views of one of the flat buffers in the parameters' own (OIHW) shapes."""
        out = []
        for p, (o, tap) in zip(self.params, self._slots):
            k = p.numel()
            if tap:
                co, ci, kh, kw = p.shape
                out.append(flat[o:o + k].view(co, kh, kw, ci).permute(0, 3, 1, 2))
            else:
                out.append(flat[o:o + k].view_as(p))
        return out

    # per-shape state: static image / GT buffers, the captured graph, its loss tensors, pinned GT staging
    class _Step:
        pass

    @property
    def graph(self):
        return self.cur.graph

    @property
    def graph_timed(self):
        return self.cur.graph_timed

    @property
    def loss(self):
        return self.cur.loss

    @property
    def log_vars(self):
        return self.cur.log_vars

    @property
    def launches_per_step(self):
        return self.cur.launches

    @staticmethod
    def _canvas(batch):
        """Shape key of a batch: [B, 3, H, W] for float images and for the uint8 [B, H, W, 3] batches of the
        device-prep pipeline (datasets.loader.collate) alike."""
        im = batch['img']
        if im.dtype == torch.uint8:
            return (im.shape[0], 3, im.shape[1], im.shape[2])
        return tuple(im.shape)

    def _stage_image(self, st, batch, src=None, hw=None):
        """Refresh the static input image of ``st``: float batches are copied; uint8 batches are uploaded as bytes and
        normalised / zero-padded by ``lsnet_image_prep_u8`` straight into the static buffer (``src`` / ``hw``: device
        staging tensors a prefetch already filled)."""
        im = batch['img']
        if im.dtype != torch.uint8:
            st.img.copy_(im if src is None else src, non_blocking=True)
            return
        if src is None:
            if getattr(st, 'u8', None) is None:
                st.u8 = torch.empty(im.shape, device=self.device, dtype=torch.uint8)
                st.hw = torch.empty((im.shape[0], 2), device=self.device, dtype=torch.int32)
            st.u8.copy_(im, non_blocking=True)
            st.hw.copy_(batch['img_hw'], non_blocking=True)
            src, hw = st.u8, st.hw
        from .datasets.loader import DevicePrep
        DevicePrep(self.device).run(src, hw, batch['img_norm_cfg'], out=st.img)

    def _gt_kwargs(self, batch):
        return dict(gt_extremes=batch.get('gt_extremes'), gt_keypoints_vs=batch.get('gt_keypoints'),
                    gt_masks=batch.get('gt_masks'))

    def _fwd_bwd(self, st):
        self.flat_g.zero_()
        losses = self.model(img=st.img, img_metas=st.metas, gt_bboxes=st.gt, gt_labels=None)
        loss, log_vars = parse_losses(losses)
        loss.backward()
        from .modules.backbone import join_fold_stream
        join_fold_stream(self.device)
        return loss, log_vars

    def _ensure_capacity(self, batch):
        """More instances than the static GT buffers hold: grow them (next power of two) and drop the captured steps (they
        are captured again on their next use)."""
        need = max(int(b.shape[0]) for b in batch['gt_bboxes'])
        if need > self.capacity:
            cap = self.capacity
            while cap < need:
                cap *= 2
            self.capacity = cap
            if self.steps:
                self.steps.clear()
                self.recaptures += 1

    def _entry(self, batch):
        self._ensure_capacity(batch)
        key = self._canvas(batch)
        st = self.steps.get(key)
        if st is None:
            if len(self.steps) >= self.max_graphs:          # drop the least recently used shape
                old = min(self.steps, key=lambda k: self.steps[k].last_use)
                del self.steps[old]
            st = self._capture(batch)
            self.steps[key] = st
        st.last_use = self.iter
        return st

    def _capture(self, batch):
        from .ops import gemm_ops
        st = GraphTrainer._Step()
        st.img = torch.empty(self._canvas(batch), device=self.device, dtype=torch.float32).contiguous(
            memory_format=torch.channels_last)
        self._stage_image(st, batch)
        with torch.no_grad():          # pyramid geometry of this input size
            feats = self.core.extract_feat(st.img[:1])
        st.sizes = [tuple(f.shape[-2:]) for f in feats]
        del feats
        st.metas = batch['img_metas']
        head = self.core.bbox_head
        st.gt = head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], st.sizes, self.device,
                             capacity=self.capacity, **self._gt_kwargs(batch))
        # warm-up on a side stream (allocator / cuDNN autotune / cudaFuncSetAttribute happen outside the capture)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._fwd_bwd(st)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gemm_ops._PACK_CACHE.clear()           # the packs must be re-done INSIDE the captured region
        from . import lib as L
        lib = L.load()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        n0 = L.launch_count()
        st.graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures
        with torch.cuda.graph(st.graph, pool=self._pool, capture_error_mode='thread_local'):
            loss, log_vars = self._fwd_bwd(st)
        # Keep only DETACHED views of the static outputs: a live grad_fn would keep this capture's autograd graph -- and its
        # AccumulateGrad nodes, bound to the streams of this capture -- alive; a later capture with another stream layout
        # (the serialised timing graph, another shape) would then find leaf streams that are not part of it.
        st.loss = loss.detach()
        st.log_vars = type(log_vars)((k, v.detach()) for k, v in log_vars.items())
        del loss, log_vars
        st.launches = L.launch_count() - n0
        gemm_ops._PACK_CACHE.clear()
        st.graph_timed = None
        if self.kernel_timing and not self.steps:
            # a second, instrumented capture of the same step (first shape only): external event-record nodes around every
            # library kernel.  Kept apart from the graph that is timed end-to-end because ~500 event nodes perturb it.
            # SERIALISED: the head's stream parallelism is switched off for this capture, so every kernel runs alone and
            # its event-bracketed duration is the kernel's own (as in an ncu launch list), not a share of a busy GPU.
            from .modules import head as _head
            saved = (_head.TOWER_STREAMS, _head.LEVEL_STREAMS, _head.REFINE_SPLIT)
            _head.TOWER_STREAMS = _head.LEVEL_STREAMS = _head.REFINE_SPLIT = False
            try:
                lib.lsnet_timing_reset()
                lib.lsnet_timing_enable(1)
                st.graph_timed = torch.cuda.CUDAGraph()
                with torch.cuda.graph(st.graph_timed, capture_error_mode='thread_local'):
                    self._fwd_bwd(st)
                lib.lsnet_timing_enable(0)
            finally:
                _head.TOWER_STREAMS, _head.LEVEL_STREAMS, _head.REFINE_SPLIT = saved
            gemm_ops._PACK_CACHE.clear()
        st.stage = None
        st.last_use = self.iter
        return st

    def replay_instrumented(self):
        """One forward+backward through the instrumented, serialised graph (gradients only; no optimizer step).
        Returns its duration in ms (CUDA events)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.graph_timed.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def load_batch(self, batch):
        """Select (or capture) the step of the batch's canvas and refresh its static inputs (host or device image tensor;
        GT lists on the host).  The packed GT goes through two persistent pinned staging sets per shape (no per-step
        cudaHostAlloc), guarded by an event each."""
        st = self.cur = self._entry(batch)
        head = self.core.bbox_head
        self._stage_image(st, batch)
        if st.stage is None:
            st.stage, st.stage_ev, st.stage_i = [], [], 0
            for _ in range(2):
                st.stage.append(head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], st.sizes, 'cpu',
                                             capacity=self.capacity, pin=True, **self._gt_kwargs(batch)))
                st.stage_ev.append(torch.cuda.Event())
        i = st.stage_i
        st.stage_i ^= 1
        st.stage_ev[i].synchronize()          # the previous async copy out of this staging set has finished
        fresh = head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], st.sizes, 'cpu',
                             capacity=self.capacity, pin=False, **self._gt_kwargs(batch))
        sg = st.stage[i]
        sg.copy_from(fresh)                   # host -> pinned host (a few KB)
        st.gt.copy_from(sg)                   # pinned host -> static device buffers, async
        st.stage_ev[i].record()

    def prefetch(self, batch):
        """Start moving the NEXT batch to the device on a copy stream while the current step computes: the image goes
        host -> a device staging buffer of its shape, the packed ground truth host -> pinned -> device staging.  The
        following ``step(batch)`` (same object) then only waits for the copy and moves staging -> static inputs on the
        device (a 50 MB device-to-device copy, ~20 us) before replaying the graph."""
        st = self._entry(batch)
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=self.device)
        u8 = batch['img'].dtype == torch.uint8
        if getattr(st, 'pre_img', None) is None or (st.pre_img.dtype == torch.uint8) != u8:
            # float batches stage the image itself, uint8 batches the bytes + extents (normalised on arrival)
            st.pre_img = torch.empty(batch['img'].shape, device=self.device, dtype=torch.uint8) if u8 else \
                torch.empty_like(st.img)
            st.pre_hw = torch.empty((batch['img'].shape[0], 2), device=self.device, dtype=torch.int32) if u8 else None
            st.pre_gt = self.core.bbox_head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], st.sizes,
                                                    self.device, capacity=self.capacity, **self._gt_kwargs(batch))
            st.pre_pin = self.core.bbox_head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], st.sizes,
                                                     'cpu', capacity=self.capacity, pin=True, **self._gt_kwargs(batch))
            st.pre_ev = torch.cuda.Event()
            st.pre_done = torch.cuda.Event()
        else:
            st.pre_done.synchronize()            # the step that consumed the previous prefetch of this shape has read it
        fresh = self.core.bbox_head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], st.sizes, 'cpu',
                                            capacity=self.capacity, pin=False, **self._gt_kwargs(batch))
        st.pre_pin.copy_from(fresh)
        cs = self._copy_stream                # (pre_done above already guarantees nobody still reads the staging buffers)
        with torch.cuda.stream(cs):
            st.pre_img.copy_(batch['img'], non_blocking=True)
            if u8:
                st.pre_hw.copy_(batch['img_hw'], non_blocking=True)
            st.pre_gt.copy_from(st.pre_pin)
            st.pre_ev.record(cs)
        st.pre_batch = batch

    def step(self, batch=None, sync_log=False, next_batch=None):
        """One training iteration.  ``next_batch``: prefetch it (host -> device on a copy stream) while this step runs."""
        if batch is not None:
            st = self.steps.get(self._canvas(batch))
            if st is not None and getattr(st, 'pre_batch', None) is batch:
                # prefetched: wait for the copy stream, staging -> static inputs on the device
                self.cur = st
                st.last_use = self.iter
                torch.cuda.current_stream(self.device).wait_event(st.pre_ev)
                self._stage_image(st, batch, src=st.pre_img, hw=st.pre_hw)
                st.gt.copy_from(st.pre_gt)
                st.pre_done.record()
                st.pre_batch = None
            else:
                self.load_batch(batch)
        self.graph.replay()
        g = self.flat_g
        if self.distributed:
            dist.all_reduce(g)
            g.div_(dist.get_world_size())
        lr = self.lr_at(self.iter)
        # clip + weight decay + momentum + update in one pass over the flat buffers (lsnet_sgd_momentum_step)
        norm = torch.linalg.vector_norm(g) if self.max_norm is not None else None
        L.call('lsnet_sgd_momentum_step', L.ptr(self.flat_p), L.ptr(g), L.ptr(self.flat_m), L.c_ll(g.numel()),
               L.ptr(norm), L.c_f(float(self.max_norm or 0.0)), L.c_f(float(lr)), L.c_f(float(self.momentum)),
               L.c_f(float(self.wd)), L.stream())
        self.iter += 1
        # the raw-pointer update does not bump the parameters' ``_version``: packs cached by an eager forward through this
        # model (evaluation, parity runs) would go stale
        from .ops import gemm_ops
        if gemm_ops._PACK_CACHE:
            gemm_ops._PACK_CACHE.clear()
        log = self.log_vars
        loss = self.loss
        if next_batch is not None:
            self.prefetch(next_batch)
        if sync_log:
            flat = torch.stack([v.detach().float() for v in log.values()])
            if self.distributed:
                dist.all_reduce(flat.div_(dist.get_world_size()))
            log = dict(zip(log.keys(), flat.tolist()))
        return loss, log


def _momentum_views(trainer):
    """{index among ALL parameters of the model (the reference builds its SGD over ``model.parameters()``, frozen ones
    included: mmcv/runner/optimizer/default_constructor.py) -> momentum buffer or None} for either trainer."""
    every = list(trainer.core.parameters())
    pos = {id(p): i for i, p in enumerate(every)}
    if hasattr(trainer, 'optimizer'):
        return {pos[id(p)]: trainer.optimizer.state.get(p, {}).get('momentum_buffer') for p in trainer.params}, every
    return {pos[id(p)]: m for p, m in zip(trainer.params, trainer.flat_views(trainer.flat_m))}, every


def save_checkpoint(trainer, path, epoch=0, meta=None):
    """The reference's checkpoint layout (mmcv/mmcv/runner/checkpoint.py:257-293, written by CheckpointHook every epoch):
    ``{'meta', 'state_dict', 'optimizer'}`` with CPU tensors, parameter names and OIHW shapes of the reference modules and
    the optimizer in ``torch.optim.SGD.state_dict()`` form over all model parameters -- so a file written here resumes
    in the reference's runner and vice versa.  ``meta`` carries ``epoch`` / ``iter`` as ``BaseRunner.resume`` expects
    (base_runner.py:289-307)."""
    state = {k: v.detach().to('cpu', copy=True).contiguous() for k, v in trainer.core.state_dict().items()}
    mom, every = _momentum_views(trainer)
    opt = dict(state={i: dict(momentum_buffer=m.detach().to('cpu', copy=True).contiguous())
                      for i, m in mom.items() if m is not None},
               param_groups=[dict(lr=float(trainer.lr_at(trainer.iter)), momentum=_sgd(trainer, 'momentum'), dampening=0,
                                  weight_decay=_sgd(trainer, 'weight_decay'), nesterov=False,
                                  params=list(range(len(every))))])
    ckpt = dict(meta=dict(meta or {}, epoch=int(epoch), iter=int(trainer.iter)), state_dict=state, optimizer=opt)
    torch.save(ckpt, path)
    return ckpt['meta']


def _sgd(trainer, key):
    if hasattr(trainer, 'optimizer'):
        return trainer.optimizer.param_groups[0][key]
    return trainer.momentum if key == 'momentum' else trainer.wd


def load_checkpoint_file(path):
    """``torch.load`` on the CPU, tensors-and-plain-containers only where the file allows it (what the reference's
    checkpoints hold: state_dict, optimizer state, a meta dict of strings / numbers); files that pickle other objects
    (e.g. a config object in ``meta``) fall back to the full unpickler, as mmcv's ``load_checkpoint`` always uses."""
    try:
        return torch.load(path, map_location='cpu', weights_only=True)
    except Exception:
        return torch.load(path, map_location='cpu', weights_only=False)


def resume(trainer, path, strict=True):
    """``BaseRunner.resume`` (mmcv/mmcv/runner/base_runner.py:289-307): weights (``module.`` prefixes of a DataParallel
    checkpoint stripped, checkpoint.py:204-240), momentum buffers, epoch and iteration.  Returns ``meta``."""
    ckpt = load_checkpoint_file(path)
    state = ckpt.get('state_dict', ckpt)
    state = {(k[7:] if k.startswith('module.') else k): v for k, v in state.items()}
    trainer.core.load_state_dict(state, strict=strict)          # copies INTO the (flat-buffer) parameter storage
    opt = ckpt.get('optimizer')
    if opt is not None:
        mom, every = _momentum_views(trainer)
        graph = not hasattr(trainer, 'optimizer')
        if graph:
            trainer.flat_m.zero_()
        for i, st in opt['state'].items():
            i = int(i)
            if i not in mom:
                raise ValueError(f'checkpoint holds a momentum buffer for parameter {i}, which is not trained here')
            buf = st['momentum_buffer']
            if graph:
                mom[i].copy_(buf)
            else:
                trainer.optimizer.state[every[i]]['momentum_buffer'] = buf.to(every[i].device).clone()
    meta = ckpt.get('meta', {})
    trainer.iter = int(meta.get('iter', 0))
    if not hasattr(trainer, 'optimizer'):
        from .ops import gemm_ops
        gemm_ops._PACK_CACHE.clear()                            # packs of the previous weights
    return meta


def train_epochs(trainer, loader, epochs=1, start_epoch=0, on_step=None):
    """The inner loops of the reference's ``EpochBasedRunner.run`` / ``train`` (mmcv/mmcv/runner/epoch_based_runner.py:
    25-44, 62-100) for a ``GraphTrainer`` (or ``Trainer``) fed by ``datasets.build_dataloader``: per epoch
    ``sampler.set_epoch`` (DistSamplerSeedHook, mmcv/runner/hooks/sampler_seed.py), then one step per batch.  With a
    GraphTrainer every step also starts the upload of the FOLLOWING batch on the copy stream (one batch of look-ahead),
    so the host-to-device copy of step i+1 overlaps the compute of step i.  ``on_step(epoch, i, loss, log_vars)`` is
    the logging hook; returns the number of steps run."""
    graph = hasattr(trainer, 'prefetch')
    n = 0
    for epoch in range(start_epoch, start_epoch + epochs):
        sampler = getattr(loader, 'sampler', None)
        if hasattr(sampler, 'set_epoch'):
            sampler.set_epoch(epoch)
        it = iter(loader)
        cur = next(it, None)
        i = 0
        while cur is not None:
            nxt = next(it, None)
            if graph:
                loss, log = trainer.step(cur, next_batch=nxt)
            else:
                b = dict(cur)
                if b['img'].dtype == torch.uint8:       # device-prep batches: normalise + pad on the GPU
                    from .datasets.loader import DevicePrep
                    b = DevicePrep(trainer.device)(b)
                    b.pop('img_hw'), b.pop('img_norm_cfg')
                else:
                    b['img'] = b['img'].to(trainer.device, non_blocking=True)
                loss, log = trainer.step(b)
            if on_step is not None:
                on_step(epoch, i, loss, log)
            cur = nxt
            i += 1
            n += 1
    return n
