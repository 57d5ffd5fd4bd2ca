from .gemm_ops import gemm, gemm_tn, conv2d_nhwc, conv2d_wgrad_nhwc, pack_conv_weight  # noqa: F401
from .conv import conv2d_same, linear_nhwc, conv2d_packed, conv2d_strided, conv2d_taps  # noqa: F401
from .dcn import (deform_conv, modulated_deform_conv, modulated_deform_conv_packed, pyramid_deform_conv,  # noqa: F401
                  dcn_im2col, dcn_col2im, dcn_forward, dcn_backward_data, dcn_backward_weight,
                  join_slices)
from .loss import (cross_iou_loss_rows, cross_iou_level_loss, sigmoid_focal_loss_sum,  # noqa: F401
                   directional_targets)
from .assign import (Pyramid, centroid_assign, atss_assign, assign_targets, pred_boxes)  # noqa: F401
from .norm import group_norm_nhwc  # noqa: F401
from .headglue import pred_reg, pred_reg_table, add_softplus, upsample_add  # noqa: F401
from .nms import nms, batched_nms, multiclass_nms_lsvr  # noqa: F401
from .stem import stem_conv, maxpool3x3s2, pack_stem_weight  # noqa: F401
