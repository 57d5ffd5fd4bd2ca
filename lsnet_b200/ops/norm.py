"""GroupNorm (+ optional ReLU) over pixel-major bf16 maps.  Round-1 implementation: torch's native group_norm in fp32
(library call) followed by a cast back to bf16 channels_last; SURVEY.md §8 row f1 replaces it with a fused kernel."""
import torch
import torch.nn.functional as F


def group_norm_nhwc(x, num_groups, weight, bias, eps=1e-5, relu=False):
    y = F.group_norm(x.float(), num_groups, weight, bias, eps)
    if relu:
        y = F.relu(y)
    return y.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
