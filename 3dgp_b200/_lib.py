"""ctypes binding of lib3dgp_b200.so -- the C ABI declared in include/gp3d_b200.h.

There is deliberately NO fallback: if the shared library is missing and cannot be built (no nvcc), every op
raises.  Nothing here imports oracle/.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib3dgp_b200.so')
_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


class ConvEpilogue(ctypes.Structure):
    """gp3d_conv_epilogue (include/gp3d_b200.h)."""
    _fields_ = [('dcoef', c_void_p), ('noise', c_void_p), ('bias', c_void_p), ('noise_per_sample', c_int), ('act', c_int), ('alpha', c_float), ('gain', c_float), ('clamp', c_float)]


class RaymarchOpts(ctypes.Structure):
    """gp3d_raymarch_opts (include/gp3d_b200.h)."""
    _fields_ = [
        ('B', c_int), ('R', c_int), ('N', c_int), ('P', c_int), ('C', c_int), ('H', c_int),
        ('ray_start', c_float), ('ray_end', c_float), ('box_half', c_float), ('noise_std', c_float),
        ('use_inf_depth', c_int), ('last_back', c_int), ('white_back_end_idx', c_int), ('clamp_mode', c_int),
        ('mlp_mode', c_int), ('seed', ctypes.c_uint64), ('offset', ctypes.c_uint64),
    ]


class ConvDesc(ctypes.Structure):
    """gp3d_conv_desc (include/gp3d_b200.h)."""
    _fields_ = [('xh', c_void_p), ('xl', c_void_p), ('wh', c_void_p), ('wl', c_void_p), ('w_format', c_int), ('x_format', c_int), ('y', c_void_p),
                ('N', c_int), ('H', c_int), ('W', c_int), ('Cin', c_int), ('Cout', c_int), ('num_slabs', c_int), ('ntaps', c_int), ('taps', c_void_p),
                ('in_stride', c_int), ('HoP', c_int), ('WoP', c_int), ('Hout', c_int), ('Wout', c_int), ('osy', c_int), ('osx', c_int), ('oy0', c_int),
                ('ox0', c_int), ('accumulate', c_int), ('epi', ctypes.POINTER(ConvEpilogue))]


class RaymarchCam(ctypes.Structure):
    """gp3d_raymarch_cam (include/gp3d_b200.h)."""
    _fields_ = [('c2w', c_void_p), ('fov', c_void_p), ('patch_scales', c_void_p), ('patch_offsets', c_void_p), ('img_h', c_int), ('img_w', c_int)]


# name -> (restype, argtypes); must list EVERY symbol of include/gp3d_b200.h (tests/test_abi.py checks this).
PROTOTYPES = {
    'gp3d_last_error': (ctypes.c_char_p, []),
    'gp3d_version': (c_int, []),
    'gp3d_built_arch': (c_int, []),
    'gp3d_bias_act': (c_int, [c_void_p] * 6 + [c_int, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_float, c_float, c_void_p]),
    'gp3d_upfirdn2d_out_size': (c_int, [c_int] * 6),
    'gp3d_upfirdn2d': (c_int, [c_void_p, c_void_p, c_void_p, c_int] + [c_int] * 4 + [c_int64] * 4 + [c_int] * 11 + [c_float]
                       + [c_int] * 2 + [c_int64] * 4 + [c_void_p]),
    'gp3d_filtered_lrelu_act': (c_int, [c_void_p, c_void_p, c_int] + [c_int] * 4 + [c_int] * 4 + [c_float] * 3 + [c_int, c_void_p]),
    'gp3d_filtered_lrelu': (c_int, [c_void_p] * 6 + [c_int] * 18 + [c_float] * 3 + [c_int] * 3 + [c_void_p]),
    'gp3d_raymarch_forward': (c_int, [c_void_p, c_int] + [c_int64] * 5 + [c_void_p] * 14 + [ctypes.POINTER(RaymarchOpts), c_void_p]),
    'gp3d_raymarch_forward_cam': (c_int, [c_void_p, c_int] + [c_int64] * 5 + [ctypes.POINTER(RaymarchCam)] + [c_void_p] * 12 + [ctypes.POINTER(RaymarchOpts), c_void_p]),
    'gp3d_generate_rays': (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p] * 3),
    'gp3d_raymarch_backward': (c_int, [c_void_p, c_int] + [c_int64] * 5 + [c_void_p] * 19 + [ctypes.POINTER(RaymarchOpts), c_void_p]),
    'gp3d_modulate': (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
    'gp3d_to_uint8': (c_int, [c_void_p] * 2 + [c_int] * 5 + [c_int64] * 4 + [c_float] * 2 + [c_void_p]),
    'gp3d_demod_act': (c_int, [c_void_p] * 3 + [c_int] + [c_void_p] * 2 + [c_int] * 6 + [c_float] * 3 + [c_void_p]),
    'gp3d_demod_act_bwd': (c_int, [c_void_p] * 5 + [c_int] + [c_void_p] * 5 + [c_int] * 4 + [c_float] * 2 + [c_void_p]),
    'gp3d_demod_act_bwd_split': (c_int, [c_void_p] * 5 + [c_int] + [c_void_p] * 4 + [c_int] + [c_void_p] * 3 + [c_int] * 4 + [c_float] * 2 + [c_void_p]),
    'gp3d_fir4_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_float] + [c_int] * 8 + [c_void_p] * 3 + [ctypes.POINTER(ConvEpilogue), c_void_p]),
    'gp3d_conv2d_nhwc_act': (c_int, [c_void_p] * 5 + [c_int] * 6 + [ctypes.POINTER(ConvEpilogue), c_void_p]),
    'gp3d_act_bwd_split': (c_int, [c_void_p] * 5 + [c_int, c_void_p] + [c_int] * 4 + [c_float] * 3 + [c_void_p]),
    'gp3d_conv2d_nhwc_bf16x3_act': (c_int, [c_void_p] * 5 + [c_int] * 6 + [ctypes.POINTER(ConvEpilogue), c_void_p]),
    'gp3d_modulate_bwd': (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    'gp3d_grad_epilogue': (c_int, [c_void_p, c_int64, c_float, c_float, c_float, c_void_p]),
    'gp3d_conv_set_wide3': (c_int, [c_int]),
    'gp3d_conv_transpose_s2_nhwc': (c_int, [c_void_p] * 4 + [c_int] * 2 + [c_void_p] + [c_int] * 5 + [c_void_p]),
    'gp3d_conv2d_nhwc_bf16': (c_int, [c_void_p] * 3 + [c_int] * 7 + [c_void_p]),
    'gp3d_conv2d_nhwc_bf16x3': (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    'gp3d_conv_nhwc': (c_int, [ctypes.POINTER(ConvDesc), c_void_p]),
    'gp3d_conv_taps_nhwc': (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p] + [c_int] * 10 + [c_void_p]),
    'gp3d_wgrad_taps_nhwc': (c_int, [c_void_p] * 5 + [c_int] * 9 + [c_void_p] + [c_int] * 4 + [c_void_p]),
    'gp3d_wgrad_taps_nhwc_fmt': (c_int, [c_void_p] * 4 + [c_int] * 2 + [c_void_p] + [c_int] * 9 + [c_void_p] + [c_int] * 4 + [c_void_p]),
    'gp3d_split_pad': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'gp3d_split_bf16': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'gp3d_adam_ema_step': (c_int, [c_void_p] * 5 + [c_int64] + [c_float] * 11 + [c_void_p, c_void_p, c_void_p]),
    'gp3d_split_bf16_pad': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}


def lib():
    """Returns the loaded CDLL, building it first if the .so is absent and nvcc is available."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)       # AttributeError => header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        assert L.gp3d_built_arch() == 100, 'lib3dgp_b200.so was not built for sm_100a'
        _lib = L
    return _lib


class Gp3dError(RuntimeError):
    pass


E_UNSUPPORTED = -2   # GP3D_E_UNSUPPORTED (include/gp3d_b200.h)
launch_count = 0   # number of successful kernel-launching C-ABI calls in this process (bench.py: gpu_launches)


def check(rc, what=''):
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = lib().gp3d_last_error().decode(errors='replace')
        raise Gp3dError(f'{what} failed (code {rc}): {msg}')


DTYPE_CODE = {'torch.float32': 0, 'torch.float16': 1, 'torch.bfloat16': 2}


def dtype_code(t):
    try:
        return DTYPE_CODE[str(t.dtype)]
    except KeyError:
        raise Gp3dError(f'unsupported dtype {t.dtype}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name='tensor'):
    if t.device.type != 'cuda':
        raise Gp3dError(f'{name} must be a CUDA tensor: 3dgp_b200 has no CPU path (device={t.device})')
