"""ORACLE — test infrastructure only.  Compiles the reference's OWN deformable-conv CUDA extension, from the sources
where they lie under /root/reference (never copied), into oracle/_ref/ for sm_100a.  It is the GPU-side oracle (the
reference's real kernels, fp32/fp64) and the "kernel to beat" in the per-kernel comparison.  Only possible in the
build container (the reference tree does not travel); the GPU box uses the prebuilt .so."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
SRC = '/root/reference/code/mmdet/ops/dcn/src'
NAME = 'deform_conv_ext'


def so_path():
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith(NAME) and f.endswith('.so'):
            return os.path.join(OUT, f)
    return None


def build(quiet=False):
    if so_path():
        return so_path()
    if not os.path.isdir(SRC):
        raise RuntimeError('reference sources not present')
    os.makedirs(OUT, exist_ok=True)
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[os.path.join(SRC, 'deform_conv_ext.cpp'), os.path.join(SRC, 'cuda', 'deform_conv_cuda.cpp'),
                             os.path.join(SRC, 'cuda', 'deform_conv_cuda_kernel.cu')],
         extra_cflags=['-DWITH_CUDA'],
         extra_cuda_cflags=['-DWITH_CUDA', '-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__',
                            '-D__CUDA_NO_HALF2_OPERATORS__'],
         build_directory=OUT, verbose=not quiet, is_python_module=False)
    return so_path()


def load_ext():
    """Import the prebuilt extension (needs torch loaded first)."""
    import importlib.util
    import torch  # noqa: F401
    p = so_path()
    if p is None:
        raise RuntimeError('oracle/_ref not built')
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build())
