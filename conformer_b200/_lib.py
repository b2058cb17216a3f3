"""ctypes binding of libconformer_b200.so (the C ABI declared in include/conformer_b200.h).

There is deliberately no fallback: if the shared library is missing the import of the compute
path fails with instructions to build it, and every compute call fails when no CUDA device is
present.  Nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libconformer_b200.so")

COMPUTE_F32, COMPUTE_BF16 = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
(KIND_ENCODER, KIND_BLOCK, KIND_FFN, KIND_MHSA, KIND_REL_ATTENTION, KIND_CONV_MODULE,
 KIND_SUBSAMPLING, KIND_DECODER) = range(8)
EPI_BIAS, EPI_BIAS_SWISH, EPI_BIAS_RELU, EPI_BIAS_RES, EPI_BIAS_F32 = range(5)


class Config(C.Structure):
    _fields_ = [("n_mel_channels", C.c_int32), ("n_blocks", C.c_int32), ("d_model", C.c_int32),
                ("n_heads", C.c_int32), ("kernel_size", C.c_int32), ("vocab_size", C.c_int32),
                ("lstm_hidden_dim", C.c_int32), ("n_lstm_layers", C.c_int32)]


class LogMelConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
                ("n_mels", C.c_int32), ("f_min", C.c_float), ("f_max", C.c_float), ("norm_slaney", C.c_int32),
                ("mel_scale_slaney", C.c_int32), ("log_floor", C.c_float)]


class NativeError(RuntimeError):
    pass


_p, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "cfm_abi_version": (C.c_int, []),
    "cfm_last_error": (C.c_char_p, []),
    "cfm_device_check": (C.c_int, []),
    "cfm_module_create": (C.c_int, [C.c_int, C.POINTER(Config), C.POINTER(_p)]),
    "cfm_module_set_param": (C.c_int, [_p, C.c_char_p, _p, _i64]),
    "cfm_module_set_grad": (C.c_int, [_p, C.c_char_p, _p, _i64]),
    "cfm_train_bind": (C.c_int, [_p, C.c_int]),
    "cfm_train_refresh": (C.c_int, [_p, _p]),
    "cfm_set_pdl": (C.c_int, [C.c_int32]),
    "cfm_train_set_graphs": (C.c_int, [_p, C.c_int32]),
    "cfm_train_graph_stats": (C.c_int, [_p, C.POINTER(C.c_int64)]),
    "cfm_encoder_train_workspace_bytes": (_i64, [_p, _i32, _i32]),
    "cfm_encoder_train_forward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _f, C.c_uint64, _p, _i64, _p]),
    "cfm_encoder_train_backward": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _i64, _p]),
    "cfm_module_finalize": (C.c_int, [_p, C.c_int, _p]),
    "cfm_module_destroy": (C.c_int, [_p]),
    "cfm_padding_mask": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "cfm_rel_pos_encoding": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "cfm_ctc_greedy_decode": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "cfm_encoder_forward_host_pipelined": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i64, _p, _p, _p]),
    "cfm_swish": (C.c_int, [_p, _p, _i64, _p]),
    "cfm_glu": (C.c_int, [_p, _p, _i64, _i32, _p]),
    "cfm_subsampled_lengths": (C.c_int, [_p, _p, _i32, _p]),
    "cfm_workspace_bytes": (_i64, [_p, _i32, _i32]),
    "cfm_ffn_forward": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_mhsa_forward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_conv_module_forward": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_block_forward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_subsampling_forward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_encoder_forward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_decoder_forward": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_encoder_forward_host": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_encoder_host_workspace_bytes": (_i64, [_p, _i32, _i32]),
    "cfm_encoder_forward_host_io": (C.c_int, [_p, _p, _i32, _p, _p, _i32, _p, _i32, _i32, _p, _i64, _p, _p, _p]),
    "cfm_encoder_host_slot_release": (C.c_int, [_p, _p]),
    "cfm_logmel_create": (C.c_int, [C.POINTER(LogMelConfig), _p, _p, C.POINTER(_p)]),
    "cfm_logmel_destroy": (C.c_int, [_p]),
    "cfm_logmel_workspace_bytes": (_i64, [_p, _i32, _i32]),
    "cfm_logmel_forward": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i64, _p]),
    "cfm_logmel_lengths": (C.c_int, [_p, _p, _i32, _i32, _p]),
    "cfm_last_launch_count": (_i64, [_p]),
    "cfm_profile_enable": (C.c_int, [_p, _i32]),
    "cfm_profile_collect": (C.c_int, [_p, C.POINTER(C.c_double), _i32]),
    "cfm_gemm_bf16": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _f, _p]),
    "cfm_gemm_bf16_tn": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p]),
    "cfm_gemm_tf32": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p]),
    "cfm_gemm_f32": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _f, _p]),
    "cfm_dwconv_glu_bn_swish": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing - build it with `python -m conformer_b200.build` "
                "(or __graft_entry__.build()); conformer_b200 has no CPU or PyTorch fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().cfm_last_error()
        raise NativeError(f"conformer_b200 native call failed (status {status}): "
                          f"{msg.decode() if msg else 'unknown error'}")
