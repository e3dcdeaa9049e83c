"""ctypes binding of libvqvae3d_b200.so (C ABI: include/vqvae3d_b200.h).

There is no CPU fallback: `lib()` raises if the nvcc-built library is missing or is not a
CUDA build, and every op wrapper in `_ops.py` refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqvae3d_b200.so")
ABI_VERSION = 13

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3

_fp = C.c_void_p   # device pointers travel as integers (tensor.data_ptr())


class ConvDesc(C.Structure):
    """struct vq3d_conv_desc"""
    _fields_ = [(n, C.c_int32) for n in
                ("B", "H", "W", "Z", "C1", "C2", "Cout", "k", "stride", "pad", "pad_circular", "pre_act", "post_act")] + \
               [(n, _fp) for n in ("x1", "x2", "w", "bias", "pre_a", "pre_b", "post_scale", "post_b", "residual", "y")]


class PreactDesc(C.Structure):
    """struct vq3d_preact_desc"""
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "W", "Z", "Cin", "Cb", "Cout", "mode")] + \
               [(n, _fp) for n in ("x", "w1", "w2", "w3", "wskip", "b1a", "b1b", "b2a", "b2b", "b3a", "b3b", "b4",
                                   "scale", "b1c", "b1d", "y", "out_w", "out_b", "out_y", "pre_w", "pre_b")]


class ConvBwd(C.Structure):
    """struct vq3d_conv_bwd"""
    _fields_ = [(n, _fp) for n in ("gy", "raw", "gx1", "gx2", "gw", "gbias", "gscalars")] + [("skip_input_grads", C.c_int32)]


# name -> (restype, argtypes); the single source of truth for tests/test_host_logic.py::test_cabi_header_and_binding_agree
SIGNATURES = {
    "vq3d_abi_version": (C.c_int, []),
    "vq3d_last_error": (C.c_char_p, []),
    "vq3d_is_cuda_build": (C.c_int, []),
    "vq3d_vq_assign": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int64, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp]),
    "vq3d_vq_assign_tc_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "vq3d_vq_assign_tc": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int64, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "vq3d_vq_loss": (C.c_int, [_fp, C.c_double, C.c_int64, _fp, _fp]),
    "vq3d_vq_ema_update": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_double, C.c_double, _fp, _fp, _fp, _fp]),
    "vq3d_vq_init_stats": (C.c_int, [_fp, C.c_int64, C.c_int, C.c_int64, _fp, _fp, _fp]),
    "vq3d_vq_init_apply": (C.c_int, [_fp, C.c_int, C.c_int, C.c_double, _fp, _fp, _fp, _fp, _fp]),
    "vq3d_vq_embed_code": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int, _fp, _fp]),
    "vq3d_vq_backward": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_double, _fp, _fp]),
    "vq3d_conv3d": (C.c_int, [C.POINTER(ConvDesc), _fp]),
    "vq3d_conv3d_tc_workspace": (C.c_size_t, [C.POINTER(ConvDesc)]),
    "vq3d_conv3d_tc": (C.c_int, [C.POINTER(ConvDesc), _fp, C.c_size_t, _fp]),
    "vq3d_conv3d_backward": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(ConvBwd), _fp]),
    "vq3d_conv1x1_backward": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(ConvBwd), _fp]),
    "vq3d_conv3d_dgrad_finish": (C.c_int, [C.POINTER(ConvDesc), _fp, _fp, _fp, _fp, _fp]),
    "vq3d_upsample2x_backward": (C.c_int, [_fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp]),
    "vq3d_huber_elu_mask_backward": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "vq3d_elu_backward": (C.c_int, [_fp, _fp, _fp, C.c_int64, _fp]),
    "vq3d_evonorm_s0_backward_sums": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int64, _fp, _fp]),
    "vq3d_evonorm_s0_backward_apply": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, _fp, _fp]),
    "vq3d_adam_amsgrad_step": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, _fp]),
    "vq3d_adam_amsgrad_step_dev": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _fp, _fp]),
    "vq3d_upsample2x": (C.c_int, [_fp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "vq3d_preact_block": (C.c_int, [C.POINTER(PreactDesc), _fp]),
    "vq3d_preact_stack": (C.c_int, [C.POINTER(PreactDesc), C.c_int, _fp, _fp]),
    "vq3d_preact_stack_thin_tc": (C.c_int, [C.POINTER(PreactDesc), C.c_int, _fp, _fp]),
    "vq3d_preact_stack_tc_workspace": (C.c_size_t, [C.POINTER(PreactDesc)]),
    "vq3d_preact_stack_tc": (C.c_int, [C.POINTER(PreactDesc), C.c_int, _fp, C.c_size_t, _fp]),
    "vq3d_preact_same_backward_workspace": (C.c_size_t, [C.POINTER(PreactDesc)]),
    "vq3d_preact_same_backward": (C.c_int, [C.POINTER(PreactDesc), _fp, _fp, C.c_size_t, _fp, _fp, _fp, _fp, _fp, _fp]),
    "vq3d_preact_up_tc_workspace": (C.c_size_t, [C.POINTER(PreactDesc)]),
    "vq3d_preact_up_tc": (C.c_int, [C.POINTER(PreactDesc), _fp, C.c_size_t, _fp]),
    "vq3d_evonorm_s0_stats": (C.c_int, [_fp, C.c_int, C.c_int64, C.c_int, C.c_double, _fp, _fp, _fp]),
    "vq3d_evonorm_s0_apply": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int64, _fp, _fp]),
    "vq3d_elu_hu_rint": (C.c_int, [_fp, C.c_int64, C.c_double, C.c_double, _fp, _fp]),
    "vq3d_elu_hu_rint_i16": (C.c_int, [_fp, C.c_int64, C.c_double, C.c_double, _fp, _fp]),
    "vq3d_hu_to_network": (C.c_int, [_fp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _fp, _fp]),
    "vq3d_huber_elu_mask": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp]),
    "vq3d_huber_elu_mask_stats": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp]),
    "vq3d_huber_elu_mask_medians_workspace": (C.c_size_t, []),
    "vq3d_huber_elu_mask_medians": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_size_t, _fp]),
}


def declare(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python 3d-vq-vae-2_b200/build.py` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        l = declare(C.CDLL(LIB_PATH))
        if l.vq3d_abi_version() != ABI_VERSION:
            raise RuntimeError(f"ABI mismatch: library {l.vq3d_abi_version()} vs binding {ABI_VERSION}")
        if l.vq3d_is_cuda_build() != 1:
            raise RuntimeError("libvqvae3d_b200.so is not a CUDA build")
        _LIB = l
    return _LIB
