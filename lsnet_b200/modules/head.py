"""LSHead on the B200 kernels — same registry name, constructor arguments, parameter names and output contract as the
reference (mmdet/models/dense_heads/lsnet_head.py:17-1437), but the whole target/loss side is batched over images and
levels on the device with no host synchronisation:

  forward      towers (DCNv2 -> GN -> ReLU), init regression, pyramid-DCN feature aggregation over 3 FPN levels,
               refine regression + classification                     (lsnet_head.py:479-755)
  loss         CentroidAssigner (init) -> predicted init boxes -> ATSSAssigner (refine) -> label scatter ->
               focal + cross-IOU per level                            (lsnet_head.py:757-1437)

Tensors between kernels are pixel-major (channels_last): bf16 feature maps, fp32 prediction maps.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..registry import HEADS, build_assigner, build_loss, build_sampler
from .dcn import ModulatedDeformConvPack, PyramidDeformConv


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class B200Conv2d(nn.Conv2d):
    """nn.Conv2d whose stride-1 'same' forward runs on the tcgen05 implicit-GEMM kernel (parameter names unchanged)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.weight._lsnet_tapmajor = True       # GraphTrainer may keep it tap-major (see train.py)

    def forward(self, x, relu=False, out_fp32=False):
        k, p, d = self.kernel_size, self.padding, self.dilation
        assert self.stride == (1, 1) and self.groups == 1 and 2 * p[0] == d[0] * (k[0] - 1) and k[0] == k[1]
        return ops.conv2d_same(x, self.weight, self.bias, padding=p[0], dilation=d[0], relu=relu, out_fp32=out_fp32)


class _ConvReLU(nn.Sequential):
    """nn.Sequential(Conv2d, ReLU) with the ReLU fused into the GEMM epilogue; keeps the '.0.weight' key."""

    def __init__(self, cin, cout):
        super().__init__(B200Conv2d(cin, cout, 1, 1, 0), nn.ReLU())

    def forward(self, x):
        return self[0](x, relu=True)


class DCNConvModule(nn.Module):
    """lsnet_head.py:1830-1849 — the GroupNorm is named ``bn`` in the reference."""

    def __init__(self, in_channels=256, out_channels=256, kernel_size=3, dilation=1, num_groups=1, dcn_pad=1):
        super().__init__()
        self.conv = ModulatedDeformConvPack(in_channels, out_channels, kernel_size, 1, dcn_pad)
        self.bn = nn.GroupNorm(num_groups, out_channels)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x, exclusive=False):
        """``exclusive``: x has no other consumer (see ModulatedDeformConvPack.forward)."""
        # the GroupNorm statistics come out of the deformable convolution's GEMM epilogue
        h = {} if (GN_EPILOGUE and x.is_cuda) else None
        # ... and the bias gradient of the convolution (= per-channel sum of the norm's dx) out of the norm's backward
        ext = BIAS_SINK and x.is_cuda and ops.norm.bias_sink_ok(self.bn.weight, self.bn.bias, self.conv.bias)
        y = self.conv(x, exclusive=exclusive, gn_holder=h, gn_groups=self.bn.num_groups, skip_bias_grad=ext)
        return ops.group_norm_nhwc(y, self.bn.num_groups, self.bn.weight, self.bn.bias, self.bn.eps, relu=True,
                                   pre_sums=h.get('sums') if h else None, bias_sink=self.conv.bias if ext else None)


class NormConvModule(nn.Module):
    """conv_module_type='norm': ConvModule(conv 3x3 -> GN -> ReLU) (lsnet_head.py:118-119)."""

    def __init__(self, cin, cout, num_groups):
        super().__init__()
        self.conv = B200Conv2d(cin, cout, 3, 1, 1, bias=False)
        self.gn = nn.GroupNorm(num_groups, cout)

    def forward(self, x, exclusive=False):
        return ops.group_norm_nhwc(self.conv(x), self.gn.num_groups, self.gn.weight, self.gn.bias, self.gn.eps, relu=True)


class PackedGT:
    """Per-batch ground truth already packed into fixed-capacity device tensors (the form LSHead.loss computes on).
    Passing one of these as ``gt_bboxes`` skips the per-step host packing, which is what makes the whole step
    capturable in a CUDA graph (static buffers refreshed with copy_ between replays).

    bbox [B,G,4] f32, count [B] i32, labels [B,G] i32, tables {'bbox': [B,G,10], 'segm': [B,G,74], 'pose': [B,G,36]},
    vs [B,G,17] or None, valid_hw [B,L,2] i32 (valid extent of every pyramid level for every image)."""

    def __init__(self, bbox, count, labels, tables, vs=None, valid_hw=None):
        self.bbox, self.count, self.labels, self.tables, self.vs, self.valid_hw = bbox, count, labels, tables, vs, valid_hw

    def copy_from(self, other):
        self.bbox.copy_(other.bbox, non_blocking=True)
        self.count.copy_(other.count, non_blocking=True)
        self.labels.copy_(other.labels, non_blocking=True)
        for k in self.tables:
            self.tables[k].copy_(other.tables[k], non_blocking=True)
        if self.vs is not None:
            self.vs.copy_(other.vs, non_blocking=True)
        if self.valid_hw is not None:
            self.valid_hw.copy_(other.valid_hw, non_blocking=True)


# GroupNorm statistics of the tower layers accumulated by the deformable convolution's GEMM epilogue
GN_EPILOGUE = os.environ.get('LSNET_GN_EPILOGUE', '1') == '1'
# bias gradient of a tower convolution added by the backward apply kernel of the GroupNorm behind it
BIAS_SINK = os.environ.get('LSNET_BIAS_SINK', '1') == '1'
TOWER_STREAMS = os.environ.get('LSNET_TOWER_STREAMS', '1') == '1'
# Every pyramid level on its own streams: the kernels of the three small levels (6 % of the pixels, 3-44 CTAs, 15-20 us
# each) run beside each other instead of one after the other.  Measured (1xB200, B=4): all levels in one chain 29.8 ms,
# '0|1,2,3,4' 29.7 ms (a persistent level-0 kernel leaves no SM for a chain that is itself serial), '0,1|2|3|4' 26.9 ms,
# '0|1|2|3|4' 25.4 ms.
LEVEL_STREAMS = os.environ.get('LSNET_LEVEL_STREAMS', '1') == '1'
# level groups for LEVEL_STREAMS, e.g. '0,1|2|3|4': one stream pair per group
LEVEL_GROUPS = os.environ.get('LSNET_LEVEL_GROUPS', '0|1|2|3|4')
# softplus / get_pred_reg / refine softplus on the library's fused element-wise kernels (torch ops otherwise)
HEAD_GLUE = os.environ.get('LSNET_HEAD_GLUE', '1') == '1'
# classification half of a level's refine stage on its own stream
REFINE_SPLIT = os.environ.get('LSNET_REFINE_SPLIT', '1') == '1'


def _level_groups(L):
    gs = [[int(t) for t in g.split(',') if int(t) < L] for g in LEVEL_GROUPS.split('|')]
    gs = [g for g in gs if g]
    if sorted(sum(gs, [])) != list(range(L)):      # a pyramid with another depth than the spec: one group per level
        gs = [[l] for l in range(L)]
    return gs
_TOWER_STREAM = {}


def _tower_stream(device, idx=0):
    key = (str(device), idx)
    if key not in _TOWER_STREAM:
        _TOWER_STREAM[key] = torch.cuda.Stream(device=device)
    return _TOWER_STREAM[key]


import contextlib
_nullctx = contextlib.nullcontext

from .losses import FocalLoss  # noqa: E402

BRANCHES = {'bbox': ['bbox'], 'segm': ['segm'], 'pose_bbox': ['bbox', 'pose'], 'pose_kbox': ['pose']}
LOSS_KIND = {'bbox': 'bbox', 'segm': 'polygon', 'pose': 'keypoint'}


@HEADS.register_module()
class LSHead(nn.Module):

    def __init__(self, num_classes, in_channels, point_feat_channels=256, num_kernel_points=9, gradient_mul=0.1,
                 point_strides=[8, 16, 32, 64, 128], point_base_scale=4, task='bbox', num_vectors=4,
                 conv_module_type='norm',
                 loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                 loss_bbox_init=dict(type='CrossIOULoss', loss_weight=1.0),
                 loss_bbox_refine=dict(type='CrossIOULoss', loss_weight=2.0), loss_segm_init=None,
                 loss_segm_refine=None, loss_pose_init=None, loss_pose_refine=None,
                 # AnchorFreeHead kwargs (anchor_free_head.py:42-61)
                 feat_channels=256, stacked_convs=4, strides=(4, 8, 16, 32, 64), dcn_on_last_conv=False,
                 conv_bias='auto', background_label=None, loss_bbox=None, conv_cfg=None, norm_cfg=None,
                 train_cfg=None, test_cfg=None):
        super().__init__()
        assert task in BRANCHES
        self.task, self.num_vectors, self.num_kernel_points = task, num_vectors, num_kernel_points
        self.num_classes, self.cls_out_channels = num_classes, num_classes
        self.in_channels, self.feat_channels, self.point_feat_channels = in_channels, feat_channels, point_feat_channels
        self.stacked_convs, self.conv_module_type, self.norm_cfg = stacked_convs, conv_module_type, norm_cfg
        self.background_label = num_classes if background_label is None else background_label
        assert self.background_label == 0 or self.background_label == num_classes   # anchor_free_head.py:79-83
        self.gradient_mul, self.point_base_scale, self.point_strides = gradient_mul, point_base_scale, list(point_strides)
        self.fpn_levels = list(range(len(self.point_strides)))
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.dcn_kernel = int(np.sqrt(num_kernel_points))
        self.dcn_pad = int((self.dcn_kernel - 1) / 2)
        assert self.dcn_kernel * self.dcn_kernel == num_kernel_points, 'The points number should be a square number.'
        assert self.dcn_kernel % 2 == 1, 'The points number should be an odd square number.'
        base = np.arange(-self.dcn_pad, self.dcn_pad + 1).astype(np.float64)
        base_offset = np.stack([np.repeat(base, self.dcn_kernel), np.tile(base, self.dcn_kernel)], axis=1).reshape(-1)
        self._base_offset_list = [float(v) for v in base_offset]
        self.register_buffer('dcn_base_offset', torch.tensor(base_offset, dtype=torch.float32).view(1, -1, 1, 1),
                             persistent=False)
        if train_cfg:
            init_cfg, refine_cfg = _cfg_get(train_cfg, 'init'), _cfg_get(train_cfg, 'refine')
            self.init_assigner = build_assigner(dict(_cfg_get(init_cfg, 'assigner')))
            self.refine_assigner = build_assigner(dict(_cfg_get(refine_cfg, 'assigner')))
            self.sampler = build_sampler(dict(type='PseudoSampler'), context=self)
        self.loss_cls = build_loss(dict(loss_cls))
        for br in BRANCHES[task]:
            cfg_i, cfg_r = {'bbox': (loss_bbox_init, loss_bbox_refine), 'segm': (loss_segm_init, loss_segm_refine),
                            'pose': (loss_pose_init, loss_pose_refine)}[br]
            setattr(self, f'loss_{br}_init', build_loss(dict(cfg_i)))
            setattr(self, f'loss_{br}_refine', build_loss(dict(cfg_r)))
        self._init_layers()

    # ------------------------------------------------------------------------------------------ layers
    def _tower(self):
        groups = self.norm_cfg['num_groups']
        mods = nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            if self.conv_module_type == 'norm':
                mods.append(NormConvModule(chn, self.feat_channels, groups))
            else:
                mods.append(DCNConvModule(chn, self.feat_channels, self.dcn_kernel, 1, groups, self.dcn_pad))
        return mods

    def _out_dims(self, br):
        if br == 'bbox':   # lsnet_head.py:170-176, 207-211: 4 extreme points + centre, plus the free DCN points
            nv = self.num_vectors if self.task == 'bbox' else 4
            return 4 * (nv + 1) + (self.num_kernel_points - nv - 1) * 2, 4 * (nv + 1)
        d = (self.num_vectors + 1) * 4
        return d, d

    def _init_layers(self):            # lsnet_head.py:93-257
        c, pc, groups = self.feat_channels, self.point_feat_channels, self.norm_cfg['num_groups']
        self.relu = nn.ReLU(inplace=True)
        self.softplus = nn.Softplus()
        self.cls_GN = nn.GroupNorm(groups, c)
        self.cls_convs = self._tower()
        for br in BRANCHES[self.task]:
            setattr(self, f'{br}_GN', nn.GroupNorm(groups, c))
            setattr(self, f'{br}_convs', self._tower())
        self.pts_cls_conv = PyramidDeformConv(c, pc, self.dcn_kernel, 1, self.dcn_pad)
        self.pts_cls_out = B200Conv2d(pc, self.cls_out_channels, 1, 1, 0)
        self.cls_af_dcn_conv = _ConvReLU(3 * pc, pc)
        self.cls_feat_conv = B200Conv2d(c, pc, 3, 1, 1)
        for br in BRANCHES[self.task]:
            d_init, d_ref = self._out_dims(br)
            setattr(self, f'pts_{br}_init_conv', B200Conv2d(c, pc, 3, 1, 1))
            setattr(self, f'pts_{br}_init_out', B200Conv2d(pc, d_init, 1, 1, 0))
            setattr(self, f'pts_{br}_refine_conv', PyramidDeformConv(c, pc, self.dcn_kernel, 1, self.dcn_pad))
            setattr(self, f'pts_{br}_refine_out', B200Conv2d(pc, d_ref, 1, 1, 0))
            setattr(self, f'{br}_af_dcn_conv', _ConvReLU(3 * pc, pc))
            setattr(self, f'{br}_feat_conv', B200Conv2d(c, pc, 3, 1, 1))

    def init_weights(self):            # lsnet_head.py:259-319
        def normal(m, std=0.01, bias=0.):
            nn.init.normal_(m.weight, 0, std)
            if getattr(m, 'bias', None) is not None:
                nn.init.constant_(m.bias, bias)

        def kaiming(m):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        towers = ['cls'] + BRANCHES[self.task]
        for t in towers:
            for m in getattr(self, f'{t}_convs'):
                normal(m.conv)
        kaiming(self.pts_cls_conv)
        normal(self.pts_cls_out, bias=float(-np.log((1 - 0.01) / 0.01)))
        normal(self.cls_feat_conv)
        normal(self.cls_af_dcn_conv[0])
        for br in BRANCHES[self.task]:
            normal(getattr(self, f'pts_{br}_init_conv'))
            normal(getattr(self, f'pts_{br}_init_out'))
            kaiming(getattr(self, f'pts_{br}_refine_conv'))
            normal(getattr(self, f'pts_{br}_refine_out'))
            normal(getattr(self, f'{br}_feat_conv'))
            normal(getattr(self, f'{br}_af_dcn_conv')[0])

    # ------------------------------------------------------------------------------------------ forward
    @staticmethod
    def _signed_pairs(t):
        """max over each (-,+) slot pair, '-' slot wins ties and is negated (lsnet_head.py:374-377)."""
        r = t.reshape(t.shape[0], -1, 2, *t.shape[2:])
        val, ind = r.max(dim=2)
        return torch.where(ind == 0, -val, val)

    def get_pred_reg(self, raw_reg1, raw_reg2):
        """lsnet_head.py:372-400: the 9 signed (y,x) sampling points handed to the pyramid DCNs."""
        if raw_reg2 is not None:
            return torch.cat((self._signed_pairs(raw_reg1), raw_reg2), dim=1)
        r = raw_reg1.reshape(raw_reg1.shape[0], -1, 4, *raw_reg1.shape[2:])
        cts, polys = r[:, -1:], r[:, :-1]
        if self.task == 'segm':
            sel = polys[:, ::math.ceil(self.num_vectors / (self.num_kernel_points - 1))]
        else:
            sel = polys[:, 1::2]
        offs = torch.cat([sel, cts], dim=1)
        offs = offs.reshape(offs.shape[0], -1, 2, *offs.shape[3:])
        val, ind = offs.max(dim=2)
        return torch.where(ind == 0, -val, val)

    _SCALE_CACHE = {}

    @staticmethod
    def _join(buf, slices):
        """The three pyramid-DCN results of a level: already adjacent in ``buf`` when they were written in place
        (the reference does torch.cat, lsnet_head.py:701,707); otherwise concatenate."""
        if all(s.data_ptr() == buf.data_ptr() + 2 * i * s.shape[1] for i, s in enumerate(slices)):
            return ops.join_slices(buf.permute(0, 3, 1, 2), slices)
        return torch.cat(slices, dim=1)

    def _scale_vec(self, sh, sw, device):
        """(1, 2*points, 1, 1) vector [sh, sw, sh, sw, ...]; cached on the device so that no host->device copy happens
        inside a captured step."""
        key = (float(sh), float(sw), self.num_kernel_points, str(device))
        v = LSHead._SCALE_CACHE.get(key)
        if v is None:
            v = torch.tensor([sh, sw] * self.num_kernel_points, dtype=torch.float32).view(1, -1, 1, 1).to(device)
            LSHead._SCALE_CACHE[key] = v
        return v

    def forward_single1(self, x, with_cls=True):
        """lsnet_head.py:502-598 for one level: towers + init regression -> (cls_feat, {br: (feat, init_sp, dcn_off)})."""
        cls_feat = x if with_cls else None
        if with_cls:
            for i, m in enumerate(self.cls_convs):
                cls_feat = m(cls_feat, exclusive=i > 0)      # layer i > 0 is the only reader of layer i-1's output
        out = {}
        for br in BRANCHES[self.task]:
            feat = x
            for i, m in enumerate(getattr(self, f'{br}_convs')):
                feat = m(feat, exclusive=i > 0)
            hid = getattr(self, f'pts_{br}_init_conv')(feat, relu=True)
            o = getattr(self, f'pts_{br}_init_out')(hid, out_fp32=True)
            if HEAD_GLUE and o.is_cuda:
                # softplus + get_pred_reg + gradient-mul mix + base offset: one kernel each way (ops/headglue.py)
                n_sp, src, mode = ops.pred_reg_table(br, self.num_vectors, self.num_kernel_points, o.shape[1])
                sp, off = ops.pred_reg(o, n_sp, src, mode, self._base_offset_list, self.gradient_mul)
                out[br] = (feat, sp, off)
                continue
            if br == 'bbox':
                sp = self.softplus(o[:, :20])
                reg = self.get_pred_reg(sp, o[:, 20:])
            else:
                sp = self.softplus(o)
                reg = self.get_pred_reg(sp, None)
            reg = (1 - self.gradient_mul) * reg.detach() + self.gradient_mul * reg
            out[br] = (feat, sp, reg - self.dcn_base_offset)
        return cls_feat, out

    def _towers_parallel(self, feats):
        """forward_single1 for every level with the classification tower on a side stream: the two towers of a level
        are independent chains, and the issue-bound gather / scatter / GroupNorm kernels of one overlap the tensor-bound
        GEMMs of the other (autograd runs each backward node on its forward stream, so the backward overlaps too).
        With LEVEL_STREAMS the small levels 1.. get their own pair of streams, so their short kernels run beside the
        level-0 ones instead of after them.  Inside the captured step these become parallel branches of the CUDA graph."""
        cur = torch.cuda.current_stream()
        dev = feats[0].device
        L = len(feats)
        groups = [list(range(L))] if not (LEVEL_STREAMS and L > 1) else _level_groups(L)
        cls_feats, outs = [None] * L, [None] * L
        used = []
        for gi, lv in enumerate(groups):
            s_cls = _tower_stream(dev, 2 * gi)
            s_reg = cur if gi == 0 else _tower_stream(dev, 2 * gi + 1)
            for st in (s_cls, s_reg):
                if st is not cur:
                    st.wait_stream(cur)
                    used.append(st)
            with torch.cuda.stream(s_cls):
                for l in lv:
                    c = feats[l]
                    for i, m in enumerate(self.cls_convs):
                        c = m(c, exclusive=i > 0)
                    cls_feats[l] = c
            with torch.cuda.stream(s_reg):
                for l in lv:
                    outs[l] = self.forward_single1(feats[l], with_cls=False)[1]
        for st in used:
            cur.wait_stream(st)
        for l in range(L):
            cls_feats[l].record_stream(cur)
            if len(groups) > 1 and l not in groups[0]:
                for br in outs[l]:
                    for t in outs[l][br]:
                        t.record_stream(cur)
        return list(zip(cls_feats, outs))

    def _refine_level(self, l, L, lvl, cls_feats, brs, cls_driver, outs):
        """lsnet_head.py:600-755 for one level: the three pyramid DCNs per branch, fusion conv + GN, refine / cls heads.
        The classification half only shares the (tiny) scaled offsets with the regression half, so with REFINE_SPLIT
        it runs on a second stream of the level."""
        lvls = [l, l + 1, l + 2] if l == 0 else ([l, l - 1, l - 2] if l == L - 1 else [l, l - 1, l + 1])
        bh, bw = cls_feats[l].shape[2:]
        dev = cls_feats[l].device
        B_, pc = cls_feats[l].shape[0], self.point_feat_channels
        # the reference scales views of the offset tensor in place, so the factors accumulate over the three
        # iterations (lsnet_head.py:628-633; SURVEY parity trap P1)
        offs = {br: [] for br in brs}
        geo = []
        for j, lv in enumerate(lvls):
            sh, sw = cls_feats[lv].size(2) / bh, cls_feats[lv].size(3) / bw
            sc = self._scale_vec(sh, sw, dev)
            geo.append((lv, sh, sw))
            for br in brs:
                prev = offs[br][-1] if j else lvl[l][1][br][2]
                offs[br].append(prev * sc)
        cur = torch.cuda.current_stream() if dev.type == 'cuda' else None
        side = None
        if cur is not None and REFINE_SPLIT and TOWER_STREAMS:
            side = _tower_stream(dev, 200 + l)
            side.wait_stream(cur)
            for t in offs[cls_driver]:
                t.record_stream(side)

        def cls_half():
            cls_buf = torch.empty((B_, bh, bw, 3 * pc), device=dev, dtype=torch.bfloat16)
            cls_raws = [self.pts_cls_conv(cls_feats[lv], offs[cls_driver][j], sh, sw, out_slice=(cls_buf, j * pc))
                        for j, (lv, sh, sw) in enumerate(geo)]
            t = ops.group_norm_nhwc(self.cls_af_dcn_conv(self._join(cls_buf, cls_raws)), self.cls_GN.num_groups,
                                    self.cls_GN.weight, self.cls_GN.bias, self.cls_GN.eps, relu=True,
                                    residual=self.cls_feat_conv(cls_feats[l]))
            return self.pts_cls_out(t, out_fp32=True)

        if side is not None:
            with torch.cuda.stream(side):
                cls_out = cls_half()
        for br in brs:
            buf = torch.empty((B_, bh, bw, 3 * pc), device=dev, dtype=torch.bfloat16)
            raws = [getattr(self, f'pts_{br}_refine_conv')(lvl[lv][1][br][0], offs[br][j], sh, sw, out_slice=(buf, j * pc))
                    for j, (lv, sh, sw) in enumerate(geo)]
            t = getattr(self, f'{br}_af_dcn_conv')(self._join(buf, raws))
            gn = getattr(self, f'{br}_GN')
            t = ops.group_norm_nhwc(t, gn.num_groups, gn.weight, gn.bias, gn.eps, relu=True,
                                    residual=getattr(self, f'{br}_feat_conv')(lvl[l][1][br][0]))
            t = getattr(self, f'pts_{br}_refine_out')(t, out_fp32=True)
            if HEAD_GLUE and t.is_cuda:
                outs[br + '_refine'].append(ops.add_softplus(t, lvl[l][1][br][1]))
            else:
                outs[br + '_refine'].append(self.softplus(t + lvl[l][1][br][1].detach()))
        if side is not None:
            cur.wait_stream(side)
            cls_out.record_stream(cur)
        else:
            cls_out = cls_half()
        outs['cls'].append(cls_out)

    def forward(self, feats):
        L = len(feats)
        if TOWER_STREAMS and feats[0].is_cuda:
            lvl = self._towers_parallel(feats)
        else:
            lvl = [self.forward_single1(f) for f in feats]
        cls_feats = [c for c, _ in lvl]
        brs = BRANCHES[self.task]
        cls_driver = brs[-1]      # pts_cls_conv follows the pose offsets when both branches exist (:680-681)
        outs = {'cls': []}
        for br in brs:
            outs[br + '_init'] = [lvl[l][1][br][1] for l in range(L)]
            outs[br + '_refine'] = []
        cur = torch.cuda.current_stream() if feats[0].is_cuda else None
        sides = {}
        if cur is not None and TOWER_STREAMS and LEVEL_STREAMS and L > 1:
            for gi, lv in enumerate(_level_groups(L)):
                if gi == 0:
                    continue
                st = _tower_stream(feats[0].device, 100 + gi)
                st.wait_stream(cur)
                for l in lv:
                    sides[l] = st
            for c, o in lvl:            # tower outputs are read by the refine stage of the neighbouring levels
                for st in set(sides.values()):
                    c.record_stream(st)
                    for br in o:
                        for t in o[br]:
                            t.record_stream(st)
        res = {}
        for l in range(L):
            lo = {k: [] for k in outs if k == 'cls' or k.endswith('_refine')}
            with torch.cuda.stream(sides.get(l, cur)) if cur is not None else _nullctx():
                self._refine_level(l, L, lvl, cls_feats, brs, cls_driver, lo)
            res[l] = lo
        for l in range(L):
            for k, v in res[l].items():
                outs[k].extend(v)
        for st in set(sides.values()):
            cur.wait_stream(st)
        for l in sides:
            for v in res[l].values():
                for t in v:
                    t.record_stream(cur)
        none = [None] * L
        return (outs['cls'], outs.get('bbox_init', none), outs.get('bbox_refine', none), outs.get('segm_init', none),
                outs.get('segm_refine', none), outs.get('pose_init', none), outs.get('pose_refine', none))

    # ------------------------------------------------------------------------------------------ loss
    def forward_train(self, x, img_metas, gt_bboxes, gt_extremes=None, gt_keypoints=None, gt_masks=None,
                      gt_labels=None, gt_bboxes_ignore=None, proposal_cfg=None, **kwargs):
        outs = self(x)
        return self.loss(*outs, gt_bboxes, gt_extremes, gt_keypoints, gt_masks, gt_labels, img_metas,
                         gt_bboxes_ignore=gt_bboxes_ignore)

    @staticmethod
    def _pack(rows, width, device, dtype=torch.float32):
        """list of per-image (G_i, width) tensors -> padded [B, Gmax, width] device tensor."""
        B, Gmax = len(rows), max(1, max(int(r.shape[0]) for r in rows))
        out = torch.zeros((B, Gmax, width), dtype=dtype)
        host = all(not r.is_cuda for r in rows)
        if host:
            for i, r in enumerate(rows):
                out[i, :r.shape[0]] = r.reshape(r.shape[0], width).to(dtype)
            return out.to(device, non_blocking=True)
        out = out.to(device)
        for i, r in enumerate(rows):
            out[i, :r.shape[0]] = r.reshape(r.shape[0], width).to(device=device, dtype=dtype)
        return out

    def get_border_center(self, gt_bboxes_list):     # lsnet_head.py:1677-1697
        res = []
        for b in gt_bboxes_list:
            x1, y1, x2, y2 = b.unbind(1)
            cx, cy = (x2 + x1) / 2.0, (y2 + y1) / 2.0
            res.append(torch.stack([cx, y1, x1, cy, cx, y2, x2, cy, cx, cy], dim=1))
        return res

    def process_polygons(self, gt_masks_list):
        """lsnet_head.py:1717-1756: per instance keep the largest polygon component, append the extent centre.
        Accepts PolygonMasks-like objects (.masks) or pre-processed (G, 2n+2) tensors."""
        polys, boxes = [], []
        for gm in gt_masks_list:
            if torch.is_tensor(gm):
                P = gm[:, :-2].reshape(gm.shape[0], -1, 2)
            else:
                # largest component per instance by shoelace area; strict '<' keeps the first maximum (:1728-1734).
                # Single-component instances (the common case) skip the areas; one tensor per image, not per instance.
                rows = []
                for comps in gm.masks:
                    best = 0
                    if len(comps) > 1:
                        areas = []
                        for q in comps:
                            q = np.asarray(q).reshape(-1, 2)
                            areas.append(0.5 * abs(float(np.dot(q[:, 0], np.roll(q[:, 1], 1)) -
                                                         np.dot(q[:, 1], np.roll(q[:, 0], 1)))))
                        for ci in range(1, len(comps)):
                            if areas[best] < areas[ci]:
                                best = ci
                    rows.append(np.asarray(comps[best], dtype=np.float64).reshape(-1, 2))
                P = torch.from_numpy(np.stack(rows).astype(np.float32))
            xmin, ymin = P[:, :, 0].min(1)[0], P[:, :, 1].min(1)[0]
            xmax, ymax = P[:, :, 0].max(1)[0], P[:, :, 1].max(1)[0]
            ct = torch.stack([(xmin + xmax) / 2, (ymin + ymax) / 2], 1).unsqueeze(1)
            polys.append(torch.cat([P, ct], dim=1).reshape(P.shape[0], -1))
            boxes.append(torch.stack([xmin, ymin, xmax, ymax], 1))
        return polys, boxes

    @staticmethod
    def process_keypoints_with_bbox(gt_bboxes_list, gt_kps_vs_list):     # lsnet_head.py:1758-1785
        kps, vss = [], []
        for b, k in zip(gt_bboxes_list, gt_kps_vs_list):
            x, y, v = k[:, 0::3], k[:, 1::3], k[:, 2::3]
            ct = torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], 1)
            kps.append(torch.cat((torch.stack((x, y), dim=2).reshape(k.size(0), -1), ct), 1))
            vss.append(v)
        return kps, vss

    @staticmethod
    def process_keypoints_with_kbox(gt_kps_vs_list):
        """lsnet_head.py:1787-1828 (task 'pose_kbox'): the ground-truth box of an instance is the extent of its VISIBLE
        keypoints (invisible ones count as +1e7 for the minimum and -1 for the maximum), its centre the landmark centre.
        Returns (keypoints+centre (G, 2n+2), key boxes (G,4), visibilities (G,n)); inputs are not modified."""
        kps, boxes, vss = [], [], []
        for k in gt_kps_vs_list:
            x, y, v = k[:, 0::3], k[:, 1::3], k[:, 2::3]
            hid = v == 0
            big, neg = torch.full_like(x, 10000000.), torch.full_like(x, -1.)
            xmin, ymin = torch.where(hid, big, x).min(1)[0], torch.where(hid, big, y).min(1)[0]
            xmax, ymax = torch.where(hid, neg, x).max(1)[0], torch.where(hid, neg, y).max(1)[0]
            ct = torch.stack([(xmin + xmax) / 2, (ymin + ymax) / 2], 1)
            kps.append(torch.cat((torch.stack((x, y), dim=2).reshape(k.size(0), -1), ct), 1))
            boxes.append(torch.stack([xmin, ymin, xmax, ymax], 1))
            vss.append(v)
        return kps, boxes, vss

    def loss(self, cls_scores, bbox_pts_preds_init, bbox_pts_preds_refine, segm_pts_preds_init, segm_pts_preds_refine,
             pose_pts_preds_init, pose_pts_preds_refine, gt_bboxes, gt_extremes, gt_keypoints_vs, gt_masks, gt_labels,
             img_metas, gt_bboxes_ignore=None, return_aux=False):
        dev = cls_scores[0].device
        task, brs = self.task, BRANCHES[self.task]
        preds = {'bbox': (bbox_pts_preds_init, bbox_pts_preds_refine), 'segm': (segm_pts_preds_init, segm_pts_preds_refine),
                 'pose': (pose_pts_preds_init, pose_pts_preds_refine)}
        if isinstance(gt_bboxes, PackedGT):
            return self._loss_packed(cls_scores, preds, gt_bboxes, img_metas, return_aux)
        tables, gt_vs = {}, None
        if task in ('bbox', 'pose_bbox'):
            if gt_extremes is None:
                gt_extremes = self.get_border_center(gt_bboxes)
            tables['bbox'] = self._pack(gt_extremes, 10, dev)
        if task == 'segm':
            polys, gt_bboxes = self.process_polygons(gt_masks)
            tables['segm'] = self._pack(polys, polys[0].shape[1], dev)
        if task in ('pose_bbox', 'pose_kbox'):
            if task == 'pose_kbox':
                kps, gt_bboxes, vss = self.process_keypoints_with_kbox(gt_keypoints_vs)
            else:
                kps, vss = self.process_keypoints_with_bbox(gt_bboxes, gt_keypoints_vs)
            tables['pose'] = self._pack(kps, kps[0].shape[1], dev)
            gt_vs = self._pack(vss, vss[0].shape[1], dev)
        gt_bb = self._pack(gt_bboxes, 4, dev)
        gt_cnt = torch.tensor([int(b.shape[0]) for b in gt_bboxes], dtype=torch.int32).to(dev, non_blocking=True)
        gt_lab = None
        if gt_labels is not None:
            gt_lab = self._pack([l.view(-1, 1) for l in gt_labels], 1, dev, torch.int32).squeeze(-1).contiguous()

        packed = PackedGT(gt_bb, gt_cnt, gt_lab, tables, gt_vs, None)
        return self._loss_packed(cls_scores, preds, packed, img_metas, return_aux)

    def pack_gt(self, gt_bboxes, gt_labels, img_metas, sizes, device, gt_extremes=None, gt_keypoints_vs=None,
                gt_masks=None, capacity=None, pin=False):
        """Host-side packing of one batch into a PackedGT on ``device`` (capacity = fixed Gmax, e.g. for CUDA graphs)."""
        task = self.task
        tables, gt_vs = {}, None

        def pack(rows, width, dtype=torch.float32):
            B = len(rows)
            G = capacity or max(1, max(int(r.shape[0]) for r in rows))
            out = torch.zeros((B, G, width), dtype=dtype)
            for i, r in enumerate(rows):
                if r.shape[0] > G:
                    raise ValueError(f'{r.shape[0]} ground-truth instances exceed the packed capacity {G}')
                out[i, :r.shape[0]] = r.reshape(r.shape[0], width).to(dtype)
            return out.pin_memory() if pin else out

        def fin(t):
            return t.pin_memory() if (pin and t is not None) else t
        if task in ('bbox', 'pose_bbox'):
            ext = gt_extremes if gt_extremes is not None else self.get_border_center(gt_bboxes)
            tables['bbox'] = pack(ext, 10)
        if task == 'segm':
            polys, gt_bboxes = self.process_polygons(gt_masks)
            tables['segm'] = pack(polys, polys[0].shape[1])
        if task in ('pose_bbox', 'pose_kbox'):
            if task == 'pose_kbox':
                kps, gt_bboxes, vss = self.process_keypoints_with_kbox(gt_keypoints_vs)
            else:
                kps, vss = self.process_keypoints_with_bbox(gt_bboxes, gt_keypoints_vs)
            tables['pose'] = pack(kps, kps[0].shape[1])
            gt_vs = pack(vss, vss[0].shape[1])
        bb = pack(gt_bboxes, 4)
        cnt = torch.tensor([int(b.shape[0]) for b in gt_bboxes], dtype=torch.int32)
        lab = pack([l.view(-1, 1) for l in gt_labels], 1, torch.int32).squeeze(-1).contiguous()
        valid = torch.tensor([[[min(int(np.ceil(m['pad_shape'][0] / s)), h), min(int(np.ceil(m['pad_shape'][1] / s)), w)]
                               for (h, w), s in zip(sizes, self.point_strides)] for m in img_metas], dtype=torch.int32)
        if str(device) == 'cpu':
            return PackedGT(bb, fin(cnt), fin(lab), tables, gt_vs, fin(valid))
        to = lambda t: None if t is None else t.to(device, non_blocking=True)
        return PackedGT(to(bb), to(cnt), to(lab), {k: to(v) for k, v in tables.items()}, to(gt_vs), to(valid))

    def _loss_packed(self, cls_scores, preds, gt, img_metas, return_aux=False):
        dev = cls_scores[0].device
        brs = BRANCHES[self.task]
        tables, gt_vs, gt_bb, gt_cnt, gt_lab = gt.tables, gt.vs, gt.bbox, gt.count, gt.labels
        sizes = [tuple(c.shape[-2:]) for c in cls_scores]
        pyr = ops.Pyramid(sizes, self.point_strides, None if gt.valid_hw is not None else
                          [m['pad_shape'][:2] for m in img_metas], dev, valid_hw=gt.valid_hw)
        init_acfg = _cfg_get(_cfg_get(self.train_cfg, 'init'), 'assigner')
        ref_acfg = _cfg_get(_cfg_get(self.train_cfg, 'refine'), 'assigner')
        # ---- init stage: nearest centre point of the GT's scale level (get_targets(stage='init')) ----
        a_init = ops.centroid_assign(pyr, gt_bb, gt_cnt, float(_cfg_get(init_acfg, 'scale', 4)))
        _, _, npos_init = ops.assign_targets(pyr, a_init, gt_lab, self.num_classes)
        # ---- refine stage: ATSS on the boxes decoded from the detached init predictions (:1333-1374) ----
        prim = brs[0]
        with torch.no_grad():
            boxes = ops.pred_boxes(pyr, [p.detach() for p in preds[prim][0]], polygon=(prim != 'bbox'))
        a_ref = ops.atss_assign(pyr, boxes, gt_bb, gt_cnt, int(_cfg_get(ref_acfg, 'topk', 9)))
        labels, lweights, npos_ref = ops.assign_targets(pyr, a_ref, gt_lab, self.num_classes)
        # num_total_pos = sum_i max(n_pos_i, 1), per GPU (lsnet_head.py:984)
        avg_init = npos_init.clamp(min=1).sum().float()
        avg_ref = npos_ref.clamp(min=1).sum().float()

        # Per-level loss terms as raw sums; the scalar arithmetic (x loss_weight / avg_factor) is applied ONCE per term on
        # the stacked per-level vector instead of level by level (the reference's per-level scalar ops are ~300 one-thread
        # kernels on the critical path between forward and backward).
        raw = {'loss_cls': []}
        scale = {}
        fast_cls = isinstance(self.loss_cls, FocalLoss) and self.loss_cls.reduction == 'mean'
        if fast_cls:
            scale['loss_cls'] = self.loss_cls.loss_weight / avg_ref
        for br in brs:
            raw[f'loss_{br}_init'], raw[f'loss_{br}_refine'] = [], []
            scale[f'loss_{br}_init'] = getattr(self, f'loss_{br}_init').loss_weight / avg_init
            scale[f'loss_{br}_refine'] = getattr(self, f'loss_{br}_refine').loss_weight / avg_ref
        B = cls_scores[0].shape[0]
        for l, s in enumerate(self.point_strides):
            off, P = int(pyr.offsets[l]), pyr.num_level[l]
            cs = cls_scores[l]
            Bc, C, H, W = cs.shape
            if not (cs.stride(1) == 1 and cs.stride(2) == W * cs.stride(3) and (Bc == 1 or cs.stride(0) == H * W * cs.stride(3))):
                cs = cs.contiguous(memory_format=torch.channels_last)      # rows below need pixel-major memory
            rows = torch.as_strided(cs, (B * H * W, C), (cs.stride(3), 1), cs.storage_offset())
            lab = labels[:, off:off + P].reshape(-1)
            lw = lweights[:, off:off + P].reshape(-1)
            if fast_cls:
                raw['loss_cls'].append(ops.sigmoid_focal_loss_sum(rows.float(), lab, lw, self.loss_cls.gamma,
                                                                  self.loss_cls.alpha))
            else:
                raw['loss_cls'].append(self.loss_cls(rows, lab, lw, avg_factor=avg_ref))
            for br in brs:
                kind = LOSS_KIND[br]
                for stage, a in (('init', a_init), ('refine', a_ref)):
                    mod = getattr(self, f'loss_{br}_{stage}')
                    pred = preds[br][0 if stage == 'init' else 1][l]
                    raw[f'loss_{br}_{stage}'].append(
                        ops.cross_iou_level_loss(pred, a, off, float(s), float(self.point_base_scale), tables[br], gt_bb,
                                                 gt_vs if kind == 'keypoint' else None, loss_type=kind, eps=mod.eps,
                                                 alpha=mod.alpha, pstride=mod.stride))
        losses = {}
        for k, v in raw.items():
            if k in scale:
                v = list((torch.stack(v) * scale[k]).unbind(0))
            losses[k] = v
        if return_aux:
            return losses, dict(assign_init=a_init, assign_refine=a_ref, labels=labels, label_weights=lweights,
                                npos_init=npos_init, npos_refine=npos_ref, boxes=boxes)
        return losses

    # ------------------------------------------------------------------------------------------ inference decode
    @staticmethod
    def _signed(pts):
        """max over each (-,+) slot pair, '-' wins ties and is negated -> (B, n, 2, H, W) as (y, x) pairs."""
        r = pts.view(pts.shape[0], -1, 2, *pts.shape[2:])
        val, ind = torch.max(r, dim=2)
        val = torch.where(ind == 0, -val, val)
        return val.view(val.shape[0], -1, 2, *val.shape[2:])

    def extreme_points2bbox(self, pts, y_first=True, extreme=False):
        """lsnet_head.py:321-347: 4 extreme points (top, left, bottom, right) -> box (left.x, top.y, right.x, bottom.y)."""
        v = self._signed(pts)
        pts_y, pts_x = (v[:, :, 0], v[:, :, 1]) if y_first else (v[:, :, 1], v[:, :, 0])
        bbox = torch.stack([pts_x[:, 1], pts_y[:, 0], pts_x[:, 3], pts_y[:, 2]], dim=1)
        if not extreme:
            return bbox
        extremes = torch.stack([pts_x[:, 0], pts_y[:, 0], pts_x[:, 1], pts_y[:, 1], pts_x[:, 2], pts_y[:, 2],
                                pts_x[:, 3], pts_y[:, 3]], dim=1)
        return extremes, bbox

    def vectors2bbox(self, pts, y_first=True, vector=False):
        """lsnet_head.py:349-370: the landmark vectors without the centre (last 4 channels) -> their extent box."""
        v = self._signed(pts[:, :-4])
        pts_y, pts_x = (v[:, :, 0], v[:, :, 1]) if y_first else (v[:, :, 1], v[:, :, 0])
        bbox = torch.stack([pts_x.min(1)[0], pts_y.min(1)[0], pts_x.max(1)[0], pts_y.max(1)[0]], 1)
        if not vector:
            return bbox
        vectors = torch.stack([pts_x, pts_y], 2).reshape(pts_y.shape[0], -1, *pts_y.shape[2:])
        return vectors, bbox

    def get_bboxes(self, cls_scores, bbox_pts_preds_init, bbox_pts_preds_refine, segm_pts_preds_init,
                   segm_pts_preds_refine, pose_pts_preds_init, pose_pts_preds_refine, img_metas, cfg=None, rescale=False,
                   nms=True):
        """lsnet_head.py:1439-1512: refine-stage landmarks -> boxes + landmark vectors per level, then per image
        ``_get_bboxes_single``.  Returns one (det_bboxes (n,5), det_pts (n, 2*num_vectors), det_labels (n,)) per image."""
        task = self.task
        f32 = lambda ts: [t.detach().float() for t in ts]
        if task in ('bbox', 'pose_bbox'):
            ext = [self.extreme_points2bbox(p, extreme=True) for p in f32(bbox_pts_preds_refine)]
        if task == 'segm':
            vec = [self.vectors2bbox(p, vector=True) for p in f32(segm_pts_preds_refine)]
        if task in ('pose_bbox', 'pose_kbox'):
            vec = [self.vectors2bbox(p, vector=True) for p in f32(pose_pts_preds_refine)]
        box_src = ext if task in ('bbox', 'pose_bbox') else vec
        pts_src = ext if task == 'bbox' else vec
        L_ = len(cls_scores)
        dev = cls_scores[0].device
        points = []
        for i in range(L_):                       # PointGenerator.grid_points (point_generator.py:17-25): (x, y, stride)
            h, w = cls_scores[i].shape[-2:]
            s = self.point_strides[i]
            xs = torch.arange(0., w, device=dev) * s
            ys = torch.arange(0., h, device=dev) * s
            points.append(torch.stack([xs.repeat(h), ys.view(-1, 1).repeat(1, w).view(-1)], -1))
        out = []
        for img_id, meta in enumerate(img_metas):
            out.append(self._get_bboxes_single([cls_scores[i][img_id].detach().float() for i in range(L_)],
                                               [box_src[i][1][img_id] for i in range(L_)],
                                               [pts_src[i][0][img_id] for i in range(L_)], points, meta['img_shape'],
                                               meta['scale_factor'], cfg, rescale, nms))
        return out

    def _get_bboxes_single(self, cls_scores, bbox_preds, pts_preds, mlvl_points, img_shape, scale_factor, cfg, rescale=False,
                           nms=True):
        """lsnet_head.py:1514-1668: top-``nms_pre`` points per level by their best class score, decode (prediction x
        stride + point), clamp to the image, multi-class NMS (``multiclass_nms_lsvr``)."""
        cfg = self.test_cfg if cfg is None else cfg
        nv = self.num_vectors
        mb, mp, ms = [], [], []
        for i, (cs, bp, pp, points) in enumerate(zip(cls_scores, bbox_preds, pts_preds, mlvl_points)):
            scores = cs.permute(1, 2, 0).reshape(-1, self.cls_out_channels).sigmoid()
            bp = bp.permute(1, 2, 0).reshape(-1, 4)
            pp = pp.permute(1, 2, 0).reshape(-1, nv * 2)
            nms_pre = _cfg_get(cfg, 'nms_pre', -1)
            if nms_pre > 0 and scores.shape[0] > nms_pre:
                _, topk = scores.max(dim=1)[0].topk(nms_pre)
                points, bp, pp, scores = points[topk], bp[topk], pp[topk], scores[topk]
            s = self.point_strides[i]
            bboxes = bp * s + torch.cat([points, points], dim=1)
            pts = pp * s + points.repeat(1, nv)
            H, W = img_shape[0], img_shape[1]
            x1, y1 = bboxes[:, 0].clamp(min=0, max=W), bboxes[:, 1].clamp(min=0, max=H)
            x2, y2 = bboxes[:, 2].clamp(min=0, max=W), bboxes[:, 3].clamp(min=0, max=H)
            mb.append(torch.stack([x1, y1, x2, y2], dim=-1))
            if self.task == 'bbox':
                # the box sides replace the matching coordinate of each extreme point (:1590-1595)
                xt, yl = pts[:, 0].clamp(min=0, max=W), pts[:, 3].clamp(min=0, max=H)
                xb, yr = pts[:, 4].clamp(min=0, max=W), pts[:, 7].clamp(min=0, max=H)
                mp.append(torch.stack([xt, y1, x1, yl, xb, y2, x2, yr], dim=-1))
            else:
                px, py = pts[:, 0::2].clamp(min=0, max=W), pts[:, 1::2].clamp(min=0, max=H)
                mp.append(torch.stack([px, py], 2).reshape(pts.size(0), -1))
            ms.append(scores)
        mb, mp, ms = torch.cat(mb), torch.cat(mp), torch.cat(ms)
        if rescale:
            sf = np.atleast_1d(np.asarray(scale_factor, dtype=np.float32))
            sf = np.tile(sf, 4)[:4] if sf.size == 1 else sf
            mb = mb / mb.new_tensor(sf)
            reps = 2 if self.task == 'bbox' else nv
            mp = mp / mp.new_tensor(np.tile(sf, 2) if self.task == 'bbox' else np.tile(sf[:2], reps))
        ms = torch.cat([ms, ms.new_zeros(ms.shape[0], 1)], dim=1)
        if not nms:
            return mb, mp, ms
        return ops.multiclass_nms_lsvr(mb, mp, ms, nv, _cfg_get(cfg, 'score_thr'), dict(_cfg_get(cfg, 'nms')),
                                       _cfg_get(cfg, 'max_per_img'))
