"""The training step of the LSNet hot path: forward + loss + backward + grad-clip + SGD(momentum), data-parallel
through torch DDP over NCCL (one process per GPU).  Mirrors what ``OptimizerHook.after_train_iter`` +
``EpochBasedRunner.train`` do per iteration (mmcv/mmcv/runner/hooks/optimizer.py:19-28,
mmcv/mmcv/runner/epoch_based_runner.py:20-47) without the reference's per-scalar all-reduce + .item() every step."""
import torch
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


class Trainer:

    def __init__(self, cfg, device='cuda', distributed=False, bucket_cap_mb=64, model=None):
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
        clip = cfg.get('grad_clip') or {}
        self.max_norm = clip.get('max_norm')
        self.iter = 0

    def step(self, batch, sync_log=False):
        """One training iteration on a per-GPU batch whose image is already on the device.  Returns the loss tensor
        (device) and log_vars (device tensors, or python floats when sync_log)."""
        for g in self.optimizer.param_groups:
            g['lr'] = warmup_lr(self.base_lr, self.iter)
        self.optimizer.zero_grad(set_to_none=True)
        losses = self.model(**batch)
        loss, log_vars = parse_losses(losses, sync_log)
        loss.backward()
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_norm, foreach=True)
        self.optimizer.step()
        self.iter += 1
        return loss, log_vars
