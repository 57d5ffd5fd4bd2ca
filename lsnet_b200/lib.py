"""ctypes binding of liblsnet_sm100.so (include/lsnet_b200.h).  PyTorch is only the allocator / stream provider:
every call passes raw device pointers and the current CUDA stream.  There is no fallback: if the library or an
sm_100 device is missing, calls raise ``LsnetError``."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get('LSNET_LIB_PATH') or os.path.join(_HERE, 'liblsnet_sm100.so')     # override: A/B builds
_lib = None
_checked_device = False

c_int, c_ll, c_f, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p


class LsnetError(RuntimeError):
    pass


class DcnDesc(ctypes.Structure):
    """``lsnet_dcn_desc`` of include/lsnet_b200.h."""
    _fields_ = [('B', c_int), ('H', c_int), ('W', c_int), ('C', c_int), ('ldx', c_ll), ('Ho', c_int), ('Wo', c_int),
                ('kh', c_int), ('kw', c_int), ('stride_h', c_int), ('stride_w', c_int), ('pad_h', c_int),
                ('pad_w', c_int), ('dil_h', c_int), ('dil_w', c_int), ('scale_h', c_f), ('scale_w', c_f),
                ('groups', c_int), ('deformable_groups', c_int), ('mask_logits', c_int), ('dtype', c_int)]


DTYPE_BF16 = 0


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise LsnetError(f'{_LIB_PATH} is missing: run `python -m lsnet_b200.build` (no CPU fallback exists)')
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.lsnet_last_error.restype = ctypes.c_char_p
        _lib.lsnet_launch_count.restype = ctypes.c_ulonglong
        for n in ('lsnet_dcn_forward_workspace_size', 'lsnet_dcn_backward_data_workspace_size',
                  'lsnet_dcn_backward_weight_workspace_size', 'lsnet_nms_workspace_size'):
            getattr(_lib, n).restype = ctypes.c_size_t
    return _lib


def require_device():
    global _checked_device
    if not _checked_device:
        lib = load()
        if not torch.cuda.is_available():
            raise LsnetError('lsnet_b200 needs a CUDA sm_100 device; there is no CPU path')
        if lib.lsnet_require_sm100() != 0:
            raise LsnetError(lib.lsnet_last_error().decode())
        _checked_device = True


def launch_count():
    return int(load().lsnet_launch_count())


def ptr(t):
    if t is None:
        return c_vp(0)
    assert t.is_cuda, 'device tensor expected'
    return c_vp(t.data_ptr())


def stream():
    return c_vp(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    require_device()
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise LsnetError(f'{name}: {lib.lsnet_last_error().decode()}')


def host_int_array(vals):
    return (c_int * len(vals))(*[int(v) for v in vals])


def host_float_array(vals):
    return (c_f * len(vals))(*[float(v) for v in vals])
