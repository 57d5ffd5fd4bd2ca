"""The reference's OWN module tests, constructor level, re-run against the B200 modules (the forward halves need a GPU:
tests/test_gpu_model.py, tests/test_gpu_trunk_conv.py): tests/test_necks.py::test_fpn :8-155,
tests/test_backbone.py::test_resnet_backbone :292-430 and ::test_resnext_backbone :654-672.  Cases outside the LSNet
path (ResNet-18/34 BasicBlock, ResNetV1d, GroupNorm trunks, plugins) are not built and are not mirrored."""
import pytest
import torch
from torch.nn.modules.batchnorm import _BatchNorm

from lsnet_b200.modules.backbone import Bottleneck, ResNet, ResNeXt
from lsnet_b200.modules.fpn import FPN


def test_fpn():
    in_channels, out_channels = [8, 16, 32, 64], 8
    with pytest.raises(AssertionError):        # num_outs != len(in_channels) - start_level
        FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, num_outs=2)
    with pytest.raises(AssertionError):        # end_level beyond the inputs
        FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, end_level=4, num_outs=2)
    with pytest.raises(AssertionError):        # num_outs != end_level - start_level
        FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, end_level=3, num_outs=1)
    with pytest.raises(AssertionError):        # unknown add_extra_convs
        FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, add_extra_convs='on_xxx', num_outs=5)
    m = FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, add_extra_convs=True, num_outs=5)
    assert m.add_extra_convs == 'on_input' and m.num_outs == 5 and len(m.lateral_convs) == 3 and len(m.fpn_convs) == 5
    m = FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, add_extra_convs=False, num_outs=5)
    assert not m.add_extra_convs and len(m.fpn_convs) == 3            # the extra levels come from pooling
    m = FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, add_extra_convs=True,
            no_norm_on_lateral=False, norm_cfg=dict(type='BN', requires_grad=True), num_outs=5)
    assert any(isinstance(x, _BatchNorm) for x in m.modules())
    for cfg in (dict(mode='bilinear', align_corners=True), dict(scale_factor=2)):
        m = FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, add_extra_convs=True, upsample_cfg=cfg,
                num_outs=5)
        assert m.add_extra_convs == 'on_input'
    for extra in ('on_input', 'on_lateral', 'on_output'):
        m = FPN(in_channels=in_channels, out_channels=out_channels, start_level=1, add_extra_convs=extra, num_outs=5)
        assert m.add_extra_convs == extra and len(m.fpn_convs) == 5


def _norm_state(model, train_state):
    return all(m.training == train_state for m in model.modules() if isinstance(m, _BatchNorm))


def test_resnet_backbone():
    with pytest.raises(KeyError):              # depth
        ResNet(20)
    with pytest.raises(AssertionError):        # 1 <= num_stages <= 4
        ResNet(50, num_stages=0)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=5)
    with pytest.raises(AssertionError):        # len(stage_with_dcn) == num_stages
        ResNet(50, dcn=dict(type='DCN', deformable_groups=1, fallback_on_stride=False), stage_with_dcn=(True,))
    with pytest.raises(AssertionError):        # len(strides) == len(dilations) == num_stages
        ResNet(50, strides=(1,), dilations=(1, 1), num_stages=3)
    with pytest.raises(TypeError):             # pretrained: a path or None
        ResNet(50).init_weights(pretrained=0)
    with pytest.raises(AssertionError):        # style
        ResNet(50, style='tensorflow')
    m = ResNet(50, norm_eval=True)
    m.init_weights()
    m.train()
    assert _norm_state(m, False)
    m = ResNet(50, frozen_stages=1)            # stem + first stage frozen
    m.init_weights()
    m.train()
    assert m.norm1.training is False
    assert all(p.requires_grad is False for layer in (m.conv1, m.norm1) for p in layer.parameters())
    assert _norm_state(m.layer1, False) and all(p.requires_grad is False for p in m.layer1.parameters())
    assert any(p.requires_grad for p in m.layer2.parameters())
    m = ResNet(50, with_cp=True)
    assert all(b.with_cp for b in m.modules() if isinstance(b, Bottleneck))
    m = ResNet(50)
    assert all(isinstance(x, _BatchNorm) for x in m.modules() if isinstance(x, (torch.nn.GroupNorm, _BatchNorm)))
    assert ResNet(50, out_indices=(0, 1, 2)).out_indices == (0, 1, 2)
    # stage widths of the four outputs (the shapes the reference test reads off a forward pass)
    assert [getattr(m, f'layer{i}')[-1].conv3.out_channels for i in range(1, 5)] == [256, 512, 1024, 2048]
    m = ResNet(50, zero_init_residual=True)
    m.init_weights()
    assert all(float(b.norm3.weight.detach().abs().sum()) == 0 for b in m.modules() if isinstance(b, Bottleneck))
    m = ResNet(50, zero_init_residual=False)
    m.init_weights()
    assert all(float(b.norm3.weight.detach().abs().sum()) > 0 for b in m.modules() if isinstance(b, Bottleneck))
    # DCN in stages 2-4 (the X-101-DCN configs): conv2 of those stages is the deformable pack, stage 1 stays plain
    m = ResNet(50, dcn=dict(type='DCN', deformable_groups=1, fallback_on_stride=False),
               stage_with_dcn=(False, True, True, True))
    assert type(m.layer1[0].conv2).__name__ == 'Conv2d'
    assert all(type(b.conv2).__name__ == 'DeformConvPack' for s in (m.layer2, m.layer3, m.layer4) for b in s)


def test_resnext_backbone():
    with pytest.raises(KeyError):
        ResNeXt(depth=18)
    m = ResNeXt(depth=50, groups=32, base_width=4)
    assert all(b.conv2.groups == 32 for b in m.modules() if isinstance(b, Bottleneck))
    assert [getattr(m, f'layer{i}')[-1].conv3.out_channels for i in range(1, 5)] == [256, 512, 1024, 2048]
    m = ResNeXt(depth=101, groups=64, base_width=4)        # the X-101-64x4d trunk of BASELINE configs 3 and 5
    assert m.layer3[0].conv2.groups == 64 and m.layer3[0].conv2.in_channels == 1024 and len(m.layer3) == 23
