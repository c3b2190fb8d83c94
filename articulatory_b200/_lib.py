"""ctypes binding of libartic_sm100.so (the C ABI declared in include/artic.h).

The product path has NO fallback: if the library is missing or a kernel reports an
error, the call raises.  Nothing here imports ``oracle``.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libartic_sm100.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2
MAX_TAPS = 48
TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16}
DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16}


class Seq(C.Structure):
    _fields_ = [("s_outer", C.c_int64), ("s_inner", C.c_int64), ("s_row", C.c_int64),
                ("n_inner", C.c_int32), ("len", C.c_int32)]


class TapConv(C.Structure):
    _fields_ = [("X", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p),
                ("res_pre", C.c_void_p), ("mask", C.c_void_p), ("res", C.c_void_p), ("res2", C.c_void_p),
                ("Y", C.c_void_p), ("Y2", C.c_void_p),
                ("x", Seq), ("y", Seq),
                ("N", C.c_int32), ("G", C.c_int32), ("Cig", C.c_int32), ("Cog", C.c_int32),
                ("q0", C.c_int32), ("nq", C.c_int32), ("si", C.c_int32), ("so", C.c_int32), ("ro", C.c_int32),
                ("ntaps", C.c_int32),
                ("off", C.c_int32 * MAX_TAPS), ("widx", C.c_int32 * MAX_TAPS),
                ("alpha", C.c_float), ("mask_slope", C.c_float), ("act_slope", C.c_float),
                ("act", C.c_int32), ("dtype", C.c_int32), ("out_dtype", C.c_int32),
                ("Wt", C.c_void_p), ("Wt_taps", C.c_int32), ("reserved_", C.c_int32),
                ("X_sp", C.c_void_p), ("Wt_sp", C.c_void_p), ("Y_sp", C.c_void_p), ("Y2_sp", C.c_void_p),
                ("x_plane", C.c_int64), ("w_plane", C.c_int64), ("y_plane", C.c_int64)]


class TapWgrad(C.Structure):
    _fields_ = [("X", C.c_void_p), ("dY", C.c_void_p), ("dW", C.c_void_p),
                ("x", Seq), ("y", Seq),
                ("N", C.c_int32), ("G", C.c_int32), ("Cig", C.c_int32), ("Cog", C.c_int32),
                ("q0", C.c_int32), ("nq", C.c_int32), ("si", C.c_int32), ("so", C.c_int32),
                ("ntaps", C.c_int32),
                ("off", C.c_int32 * MAX_TAPS), ("yoff", C.c_int32 * MAX_TAPS), ("widx", C.c_int32 * MAX_TAPS),
                ("dtype", C.c_int32), ("y_dtype", C.c_int32),
                ("X_sp", C.c_void_p), ("dY_sp", C.c_void_p), ("x_plane", C.c_int64), ("y_plane", C.c_int64),
                ("dbias", C.c_void_p)]


class ResUnit(C.Structure):
    _fields_ = [("AX", C.c_void_p), ("XRES", C.c_void_p), ("W1t", C.c_void_p), ("W2t", C.c_void_p),
                ("b1", C.c_void_p), ("b2", C.c_void_p), ("AT", C.c_void_p), ("Y", C.c_void_p), ("Y2", C.c_void_p),
                ("N", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("k", C.c_int32), ("dil", C.c_int32),
                ("slope", C.c_float), ("mode", C.c_int32), ("reserved_", C.c_int32), ("M1", C.c_void_p), ("M2", C.c_void_p)]


class Mlp(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("W", C.c_void_p * 8), ("bias", C.c_void_p * 8), ("act0", C.c_void_p),
                ("outs", C.c_void_p * 8), ("dims", C.c_int32 * 9), ("B", C.c_int32), ("n_layers", C.c_int32),
                ("dtype", C.c_int32), ("slope", C.c_float)]


class DiscPrep(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ar", C.c_void_p), ("y", C.c_void_p * 2), ("x_out", C.c_void_p),
                ("pool", C.c_void_p * 4), ("xp", C.c_void_p * 8), ("pool_len", C.c_int32 * 4), ("xp_len", C.c_int32 * 8),
                ("N", C.c_int32), ("B", C.c_int32), ("T", C.c_int32), ("La", C.c_int32), ("n_pool", C.c_int32),
                ("n_xp", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("reserved_", C.c_int32)]


class AdamHyper(C.Structure):
    _fields_ = [("lr0", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("gamma", C.c_float), ("step", C.c_int32), ("n_milestones", C.c_int32),
                ("milestones", C.c_int32 * 8)]


class WDesc(C.Structure):
    _fields_ = [("v", C.c_void_p), ("g", C.c_void_p), ("scale", C.c_void_p),
                ("out_f", C.c_void_p), ("out_b", C.c_void_p),
                ("dWp", C.c_void_p), ("dv", C.c_void_p), ("dg", C.c_void_p),
                ("row_len", C.c_int64), ("sk", C.c_int64), ("sg", C.c_int64), ("sa", C.c_int64), ("sb", C.c_int64),
                ("tile_begin", C.c_int64), ("tile2_begin", C.c_int64),
                ("rows", C.c_int32), ("K", C.c_int32), ("G", C.c_int32), ("A", C.c_int32), ("B", C.c_int32),
                ("merge", C.c_int32), ("a_pad", C.c_int32), ("b_pad", C.c_int32),
                ("dtype_f", C.c_int32), ("dtype_b", C.c_int32), ("dw_swapped", C.c_int32), ("row_chunk", C.c_int32)]


def upload_structs(items, device):
    """Array of ctypes structs -> device uint8 tensor (descriptor tables of the batched kernels)."""
    arr = (type(items[0]) * len(items))(*items)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


_i32, _i64, _f, _p = C.c_int32, C.c_int64, C.c_float, C.c_void_p

#: every exported symbol of include/artic.h: name -> (restype, argtypes)
SIGNATURES = {
    "artic_version": (C.c_int, []),
    "artic_arch": (C.c_char_p, []),
    "artic_last_error": (C.c_char_p, []),
    "artic_debug_set": (C.c_int, [C.c_int, C.c_int]),
    "artic_debug_get": (C.c_int, [C.c_int]),
    "artic_debug_buffer": (C.c_int, [_p]),
    "artic_trace_buffer": (C.c_int, [_p, C.c_longlong]),
    "artic_tapconv": (C.c_int, [C.POINTER(TapConv), _p]),
    "artic_tapconv_multi": (C.c_int, [C.POINTER(TapConv), _i32, _p]),
    "artic_disc_prep": (C.c_int, [C.POINTER(DiscPrep), _p]),
    "artic_mlp_fwd": (C.c_int, [C.POINTER(Mlp), _p]),
    "artic_resunit_fwd": (C.c_int, [C.POINTER(ResUnit), _p]),
    "artic_tapconv_wgrad": (C.c_int, [C.POINTER(TapWgrad), _p]),
    "artic_colsum": (C.c_int, [_p, C.POINTER(Seq), _i32, _i32, _i32, _p, _p]),
    "artic_wperm_tiles": (C.c_int64, [_i32, _i32, _i32, _i32]),
    "artic_wrow_tiles": (C.c_int64, [_i32, _i32, _i32, _i32, _i64, _i64, _i64, _i32]),
    "artic_weights_prep": (C.c_int, [_p, _i32, _i32, _i64, _i64, _p]),
    "artic_weights_unprep": (C.c_int, [_p, _i32, _i32, _i64, _i64, _p]),
    "artic_gen_input": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "artic_gen_input_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "artic_mean3_act": (C.c_int, [_p, _p, _p, _p, _i64, _f, _i32, _i32, _p]),
    "artic_sum3": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p]),
    "artic_tanh_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "artic_cast": (C.c_int, [_p, _i32, _p, _i32, _i64, _p]),
    "artic_cut_windows": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "artic_split": (C.c_int, [_p, _p, _i64, _i64, _p]),
    "artic_path_counts": (C.c_int, [C.POINTER(C.c_int64), _i32]),
    "artic_concat_time": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i64, _i32, _p]),
    "artic_avgpool1d": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "artic_avgpool1d_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "artic_reflect_pad_right": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p]),
    "artic_reflect_pad_right_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "artic_sqerr_sum": (C.c_int, [_p, _i64, _f, _f, _p, _i32, _p]),
    "artic_sqerr_bwd": (C.c_int, [_p, _i64, _f, _f, _p, _i32, _i32, _p]),
    "artic_l1_sum": (C.c_int, [_p, _p, _i64, _f, _p, _i32, _p]),
    "artic_l1_sum_bwd": (C.c_int, [_p, _p, _i64, _f, _p, _f, _p, _i32, _p]),
    "artic_l1_bwd": (C.c_int, [_p, _p, _i64, _f, _p, _i32, _i32, _p]),
    "artic_stft_loss_fwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p, _f, _p, _p]),
    "artic_stft_loss_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p, _f, _p, _f, _f, _p, _p]),
    "artic_mel_loss_fwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _f, _f, _f, _p, _p]),
    "artic_mel_loss_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _f, _f, _f, _p, _p]),
    "artic_mel_loss_fwd_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _f, _f, _f, _p, _f, _p, _p]),
    "artic_add_rows": (C.c_int, [_p, _i64, _p, _i64, _i32, _i32, _p]),
    "artic_train_log": (C.c_int, [_p, _p, _p, _i32, _f, _f, _f, _p, _p, _p]),
    "artic_bigru_layer": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p]),
    "artic_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p]),
    "artic_adam_step_wire": (C.c_int, [_p, _p, _i32, _p, _p, _i64, _p, _p]),
    "artic_adam_tick": (C.c_int, [_p, _p]),
}

_lib = None
#: number of kernel-launching C-ABI calls issued by this process (bench.py reports it)
launch_count = 0


class ArticError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ArticError(
            f"{LIB_PATH} not found: build it with `python -m articulatory_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # tuning / debugging knobs of the tensor-core path: ARTIC_TC=0 forces the CUDA-core kernel,
    # ARTIC_DEBUG="k=v,k=v" sets raw artic_debug_set keys
    if os.environ.get("ARTIC_TC", "1") == "0":
        lib.artic_debug_set(1, 1)
    for kv in filter(None, os.environ.get("ARTIC_DEBUG", "").split(",")):
        k, v = kv.split("=")
        lib.artic_debug_set(int(k), int(v))
        ENV_DEBUG[int(k)] = int(v)
    _lib = lib
    return lib


#: raw debug keys set through ARTIC_DEBUG (scoped overrides restore these values)
ENV_DEBUG = {}


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Call a kernel-launching entry point on the current CUDA stream; raise on error."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args, stream_ptr())
    launch_count += 1
    if rc != 0:
        raise ArticError(f"{name} failed ({rc}): {lib.artic_last_error().decode()}")


PATH_NAMES = ("conv_tc", "conv_tc_x3", "conv_generic", "conv_c1", "wgrad_tc", "wgrad_tc_x3", "wgrad_generic", "wgrad_c1",
              "wgrad_bias_fused", "conv_tc_cluster", "conv_tc_fused")


def path_counts(reset=False):
    """Which kernel family took each contraction since the last reset (artic_path_counts)."""
    buf = (C.c_int64 * 12)()
    load().artic_path_counts(buf, int(reset))
    return dict(zip(PATH_NAMES, list(buf)[:len(PATH_NAMES)]))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise ArticError(f"{what} must live on a CUDA device: the articulatory_b200 hot path is "
                         "sm_100a kernels only (no CPU fallback)")
