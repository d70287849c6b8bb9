"""ctypes binding of ``libsrnn_b200.so`` (the C ABI declared in ``include/srnn_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  torch is used only to own device memory and streams.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libsrnn_b200.so')

_lib = None


class GemmArgs(C.Structure):
    _fields_ = [
        ('op', C.c_int32), ('m', C.c_int32), ('n', C.c_int32), ('k', C.c_int32), ('batch', C.c_int32),
        ('a', C.c_void_p), ('lda', C.c_int64), ('a_batch_stride', C.c_int64), ('a_row_offset', C.c_int32),
        ('b', C.c_void_p), ('ldb', C.c_int64), ('b_batch_stride', C.c_int64), ('b_row_offset', C.c_int32),
        ('c', C.c_void_p), ('ldc', C.c_int64), ('c_batch_stride', C.c_int64),
        ('c_dtype', C.c_int32), ('n_fold', C.c_int32),
        ('bias', C.c_void_p),
        ('aux', C.c_void_p), ('ldaux', C.c_int64), ('aux_batch_stride', C.c_int64),
        ('aux_mode', C.c_int32), ('relu', C.c_int32), ('aux_row_div', C.c_int32), ('max_ctas', C.c_int32),
        ('colsum', C.c_void_p),
        ('a2', C.c_void_p), ('lda2', C.c_int64), ('a2_batch_stride', C.c_int64), ('k1', C.c_int32),
        ('relu_mask', C.c_void_p), ('gate_mask', C.c_void_p), ('ldmask', C.c_int64),
    ]


class NllArgs(C.Structure):
    _fields_ = [
        ('mode', C.c_int32), ('m', C.c_int32), ('k', C.c_int32),
        ('a', C.c_void_p), ('lda', C.c_int64),
        ('w', C.c_void_p), ('ldw', C.c_int64),
        ('bias', C.c_void_p), ('target', C.c_void_p),
        ('lse', C.c_void_p), ('logp_target', C.c_void_p),
        ('logp', C.c_void_p), ('ldlogp', C.c_int64),
        ('row_grad', C.c_void_p),
        ('g', C.c_void_p), ('ldg', C.c_int64),
        ('dlogits', C.c_void_p), ('lddlogits', C.c_int64),
    ]


class GruArgs(C.Structure):
    _fields_ = [
        ('batch', C.c_int32), ('steps', C.c_int32), ('hidden', C.c_int32), ('ext_batch', C.c_int32),
        ('gi', C.c_void_p), ('w_hh', C.c_void_p), ('b_hh', C.c_void_p),
        ('h_ext', C.c_void_p), ('hall', C.c_void_p), ('h_state', C.c_void_p), ('gates', C.c_void_p),
        ('dh_out', C.c_void_p), ('dgi', C.c_void_p), ('dgh', C.c_void_p), ('dh0', C.c_void_p),
        ('sync', C.c_void_p), ('tuning_flags', C.c_int32), ('debug_ts', C.c_void_p),
        ('cell', C.c_int32), ('c_state', C.c_void_p), ('c_init', C.c_void_p), ('dc0', C.c_void_p),
        ('units_per_cta', C.c_int32), ('db_ih', C.c_void_p), ('db_hh', C.c_void_p),
    ]


class GruF32Args(C.Structure):
    _fields_ = [
        ('batch', C.c_int32), ('steps', C.c_int32), ('hidden', C.c_int32),
        ('gi', C.c_void_p), ('w3', C.c_void_p), ('b_hh', C.c_void_p), ('h_state', C.c_void_p), ('hall', C.c_void_p),
        ('h_init', C.c_void_p), ('gates', C.c_void_p), ('a3', C.c_void_p), ('ws', C.c_void_p),
        ('dh_out', C.c_void_p), ('dgi', C.c_void_p), ('dgh', C.c_void_p), ('dh0', C.c_void_p), ('carry', C.c_void_p),
    ]


P, I32, I64, F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> argument ctypes (every function returns int unless noted); must mirror include/srnn_b200.h
SIGNATURES = {
    'srnn_abi_version': [],
    'srnn_device_info': [P, P, P],
    'srnn_probe_clusters': [I32, I32, I32, P],
    'srnn_quantize_ulaw': [P, I64, P, P, P, P],
    'srnn_quantize_linear': [P, I64, I64, I32, P, P, P],
    'srnn_dequantize_lut': [P, P, I64, P, P, P, P],
    'srnn_onehot_rows': [P, I64, I32, P, P],
    'srnn_weight_prep': [P, P, I32, I32, I32, P, P, P, P, P, P],
    'srnn_weight_prep_bwd': [P, P, P, P, P, I32, I32, I32, P, P, P],
    'srnn_pad_cast_bf16': [P, I64, I32, I64, P, I32, I64, P],
    'srnn_split_bf16': [P, I64, I32, I64, P, P, I32, I64, P],
    'srnn_bf16_to_f32': [P, I64, I32, I64, P, I64, I32, P],
    'srnn_mixer_input': [P, P, P, I32, I32, I32, I32, P, I32, P],
    'srnn_mixer_input_bwd': [P, P, I32, I32, I32, I32, P, P],
    'srnn_tier_input': [P, I64, I32, P, P, P, I32, I32, I32, I32, I32, P, I32, P],
    'srnn_tier_input_bwd': [P, I32, I32, I32, I32, I32, I32, P, P],
    'srnn_repeat_rows': [P, I64, I32, I64, I32, P, I64, P],
    'srnn_repeat_rows_bwd': [P, I64, I32, I64, I32, P, I64, P],
    'srnn_colsum': [P, I64, I32, I64, P, P],
    'srnn_set_pdl': [I32],
    'srnn_embed_sum': [P, P, I64, I32, I32, I32, I32, P, I64, I32, P, I64, P],
    'srnn_sample_categorical': [P, I64, I32, I32, I32, P, I64, P, P, P, I32, P, I64, P],
    'srnn_sample_embed': [P, I64, I32, I32, I32, P, I64, P, P, P, I32, P, I64, P, I32, I32, P, I64, P, I64, P],
    'srnn_gemm_bf16': [C.POINTER(GemmArgs), P],
    'srnn_gemm_nll': [C.POINTER(NllArgs), P],
    'srnn_gru_forward': [C.POINTER(GruArgs), P],
    'srnn_gru_backward': [C.POINTER(GruArgs), P],
    'srnn_state_select': [P, P, P, I32, I32, P, P, I64, P],
    'srnn_state_select_bwd': [P, P, I32, I32, P, P],
    'srnn_masked_nll_mean': [P, P, I64, I32, P, P],
    'srnn_adam_clipped': [P, P, P, P, I64, F64, F64, F64, F64, I32, F64, P],
    # fp32-tolerance mode (csrc/precise.cu)
    'srnn_split3_bf16': [P, I64, I32, I64, P, I32, I64, I32, P],
    'srnn_mixer_input_f32': [P, P, P, I32, I32, I32, I32, P, I32, P],
    'srnn_mixer_input_bwd_f32': [P, P, I32, I32, I32, I32, P, P],
    'srnn_tier_input_f32': [P, I64, I32, P, P, P, I32, I32, I32, I32, I32, P, I32, P],
    'srnn_tier_input_bwd_f32': [P, I32, I32, I32, I32, I32, I64, P, P],
    'srnn_weight_prep_f32': [P, P, I32, I32, I32, P, P, P, P, P, P],
    'srnn_bias_act_f32': [P, I64, I32, I64, P, I64, I32, P, I64, I32, P, I64, P],
    'srnn_embed_gather_f32': [P, P, I64, I32, I32, I32, I32, I32, P, P],
    'srnn_segment_sum_f32': [P, I64, I32, I64, I32, P, I64, P],
    'srnn_colsum_f32': [P, I64, I32, I64, P, P],
    'srnn_logsoftmax_nll_f32': [P, I64, I64, I32, P, P, P, P],
    'srnn_logsoftmax_nll_bwd_f32': [P, I64, I64, I32, P, P, P, I64, P, I64, P],
    'srnn_gru_forward_f32': [C.POINTER(GruF32Args), P],
    'srnn_gru_backward_f32': [C.POINTER(GruF32Args), P],
}

EXPORTS = sorted(list(SIGNATURES) + ['srnn_last_error'])


def load(path=None):
    """Load (once) and return the ctypes handle.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f'{path} is missing: build it with `python -m samplernn_pase_b200._build` '
            '(nvcc, sm_100a).  samplernn_pase_b200 has no CPU or eager fallback.')
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.srnn_last_error.argtypes = []
    lib.srnn_last_error.restype = C.c_char_p
    if lib.srnn_abi_version() != 6:
        raise RuntimeError('libsrnn_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().srnn_last_error().decode(errors='replace')
        raise RuntimeError(f'{what} failed with code {rc}: {msg}')


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Raw device pointer of a tensor (or NULL).  The tensor must be a CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('samplernn_pase_b200 kernels need CUDA tensors (there is no CPU fallback)')
    return C.c_void_p(t.data_ptr())


#: when a list, every C call is bracketed with CUDA events: (name, note, start, end) tuples are appended
#: (scripts/profile_step.py); None in normal operation.
profile_log = None
profile_note = ''


def call(name, *args):
    lib = load()
    if profile_log is None:
        check(getattr(lib, name)(*args), name)
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    check(getattr(lib, name)(*args), name)
    end.record()
    profile_log.append((name, profile_note, start, end))


def device_info():
    sms, major, minor = C.c_int(0), C.c_int(0), C.c_int(0)
    call('srnn_device_info', C.byref(sms), C.byref(major), C.byref(minor))
    return sms.value, major.value, minor.value
