"""Tensor-level wrappers of the C ABI: they validate dtype/device/contiguity, pass raw pointers
and leading dimensions, and allocate outputs with torch (device memory plumbing only).

All matrices are row-major.  ``bf16`` operands must have 16-byte aligned bases and leading
dimensions that are multiples of 8 elements (TMA).
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import GemmArgs, GruArgs, NllArgs, call, ptr, stream

BF16 = torch.bfloat16
F32 = torch.float32

#: launches issued through this module since the last reset (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n=1):
    global launch_count
    launch_count += n


#: > 0 caps the persistent GEMM grids (set while GEMMs share the GPU with a recurrent kernel on another stream)
gemm_max_ctas = 0

#: when a dict, ``timed(tag)`` brackets the enclosed launches with CUDA events on the current stream
#: and appends (start, end) to ``event_log[tag]``; bench.py reads them after the timed region.
event_log = None


class timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if event_log is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if event_log is not None:
            self.end.record()
            event_log.setdefault(self.tag, []).append((self.start, self.end))
        return False


def round_up(x, m):
    return (x + m - 1) // m * m


def _need(t, dtype, name):
    if t.dtype != dtype or not t.is_cuda:
        raise RuntimeError(f'{name}: expected a CUDA {dtype} tensor, got {t.dtype} on {t.device}')


def _strides(s):
    return (C.c_int64 * 3)(*s)


# ----------------------------------------------------------------------------------------------
# quantiser
# ----------------------------------------------------------------------------------------------
def quantize_ulaw(x, want_i64=True, want_u8=False, overflow=None):
    _need(x, F32, 'quantize_ulaw x')
    x = x.contiguous()
    o64 = torch.empty(x.shape, dtype=torch.int64, device=x.device) if want_i64 else None
    o8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_u8 else None
    call('srnn_quantize_ulaw', ptr(x), x.numel(), ptr(o64), ptr(o8), ptr(overflow), stream())
    _count()
    return o64, o8


def quantize_linear(x, want_i64=True, want_u8=False, q_levels=256):
    _need(x, F32, 'quantize_linear x')
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    o64 = torch.empty(x.shape, dtype=torch.int64, device=x.device) if want_i64 else None
    o8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_u8 else None
    call('srnn_quantize_linear', ptr(x), rows, cols, int(q_levels), ptr(o64), ptr(o8), stream())
    _count()
    return o64, o8


def dequantize_lut(idx, lut, out_dtype=F32):
    _need(lut, F32, 'dequantize lut')
    idx = idx.contiguous()
    out = torch.empty(idx.shape, dtype=out_dtype, device=idx.device)
    i64 = idx if idx.dtype == torch.int64 else None
    i8 = idx if idx.dtype == torch.uint8 else None
    if i64 is None and i8 is None:
        raise RuntimeError('dequantize_lut: indices must be int64 or uint8')
    call('srnn_dequantize_lut', ptr(i64), ptr(i8), idx.numel(), ptr(lut), ptr(out if out_dtype == F32 else None),
         ptr(out if out_dtype == BF16 else None), stream())
    _count()
    return out


def onehot_rows(idx_u8, q=256):
    _need(idx_u8, torch.uint8, 'onehot idx')
    idx_u8 = idx_u8.contiguous()
    out = torch.empty(idx_u8.shape + (q,), dtype=BF16, device=idx_u8.device)
    call('srnn_onehot_rows', ptr(idx_u8), idx_u8.numel(), q, ptr(out), stream())
    _count()
    return out


def set_pdl(on):
    """Programmatic dependent launch for the per-sample generation kernels (process-wide switch)."""
    call('srnn_set_pdl', int(bool(on)))


def embed_sum(table, idx_u8, idx_ld, batch, r0, q, hidden, pre, pre_ld, relu, out, out_ld):
    """out[b] = act(sum_k table[k*q + idx[b*idx_ld + k]] + pre[b]); ``idx_u8`` may be a view into a wider window."""
    _need(table, BF16, 'embed_sum table')
    _need(idx_u8, torch.uint8, 'embed_sum idx')
    call('srnn_embed_sum', ptr(table), ptr(idx_u8), idx_ld, batch, r0, q, hidden, ptr(pre), pre_ld, int(relu), ptr(out),
         out_ld, stream())
    _count()
    return out


def sample_categorical(x, batch, q, u, win, win_len, out, out_ld, normalise=False, logp_out=None, rng_state=None):
    """Draw one code per row from exp(log-probabilities) (inverse CDF with the uniforms ``u``; arg-max when ``u`` is
    None), append it to the row's window ``win`` (shifted left by one) and store it at ``out[b*out_ld]``.  With
    ``normalise`` the rows of ``x`` are raw logits and the log-softmax is taken first (written to ``logp_out``)."""
    _need(x, F32, 'sample input')
    call('srnn_sample_categorical', ptr(x), x.stride(0), batch, q, int(normalise), ptr(logp_out),
         logp_out.stride(0) if logp_out is not None else 0, ptr(u), ptr(rng_state), ptr(win), win_len, ptr(out), out_ld,
         stream())
    _count()


def sample_embed(x, batch, q, win, win_len, out, out_ld, table, r0, hidden, pre_next, pre_ld, h1_next, h1_ld,
                 normalise=False, logp_out=None, rng_state=None, u=None):
    """``sample_categorical`` + the embedding head of the NEXT sample step in one launch (see srnn_sample_embed)."""
    _need(x, F32, 'sample input')
    call('srnn_sample_embed', ptr(x), x.stride(0), batch, q, int(normalise), ptr(logp_out),
         logp_out.stride(0) if logp_out is not None else 0, ptr(u), ptr(rng_state), ptr(win), win_len, ptr(out), out_ld,
         ptr(table), r0, hidden, ptr(pre_next), pre_ld, ptr(h1_next), h1_ld, stream())
    _count()


# ----------------------------------------------------------------------------------------------
# parameter preparation
# ----------------------------------------------------------------------------------------------
def weight_prep(v, g, shape3, out1, s1, out2=None, s2=None, inv_norm=None):
    """v viewed as (R,A,B) -> bf16 GEMM layouts (weight-normed when g is given)."""
    _need(v, F32, 'weight_prep v')
    r, a, b = shape3
    assert v.is_contiguous() and v.numel() == r * a * b
    call('srnn_weight_prep', ptr(v), ptr(g), r, a, b, ptr(out1), _strides(s1), ptr(out2),
         _strides(s2) if s2 is not None else None, ptr(inv_norm), stream())
    _count()


def weight_prep_bwd(dw, s, v, g, inv_norm, shape3):
    r, a, b = shape3
    _need(dw, F32, 'weight_prep_bwd dw')
    dv = torch.empty_like(v)
    dg = torch.empty(r, dtype=F32, device=v.device) if g is not None else None
    call('srnn_weight_prep_bwd', ptr(dw), _strides(s), ptr(v), ptr(g), ptr(inv_norm), r, a, b, ptr(dv), ptr(dg), stream())
    _count()
    return dv, dg


def pad_cast_bf16(x, rows, cols, ld_in, out, cols_pad, ld_out):
    _need(x, F32, 'pad_cast x')
    call('srnn_pad_cast_bf16', ptr(x), rows, cols, ld_in, ptr(out), cols_pad, ld_out, stream())
    _count()


def to_bf16(x):
    """fp32 (rows, cols) contiguous -> new bf16 tensor of the same shape."""
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    pad_cast_bf16(x, rows, cols, cols, out, cols, cols)
    return out


def split_bf16(x):
    """fp32 (rows, cols) contiguous -> (hi, lo) bf16 tensors with x = hi + lo to 16 mantissa bits."""
    x = x.contiguous()
    _need(x, F32, 'split_bf16 x')
    cols = x.shape[-1]
    rows = x.numel() // cols
    hi = torch.empty(x.shape, dtype=BF16, device=x.device)
    lo = torch.empty(x.shape, dtype=BF16, device=x.device)
    call('srnn_split_bf16', ptr(x), rows, cols, cols, ptr(hi), ptr(lo), cols, cols, stream())
    _count()
    return hi, lo


def bf16_to_f32(x, rows, cols, ld_in, out, ld_out, accumulate=False):
    call('srnn_bf16_to_f32', ptr(x), rows, cols, ld_in, ptr(out), ld_out, int(accumulate), stream())
    _count()


def colsum(x, rows, cols, ld):
    out = torch.empty(cols, dtype=F32, device=x.device)
    call('srnn_colsum', ptr(x), rows, cols, ld, ptr(out), stream())
    _count(2)
    return out


# ----------------------------------------------------------------------------------------------
# operand assembly
# ----------------------------------------------------------------------------------------------
def mixer_input(utt, table, spk_ids, k_pad):
    b, l, u = utt.shape
    s = table.shape[1]
    out = torch.empty(b * l, k_pad, dtype=BF16, device=utt.device)
    call('srnn_mixer_input', ptr(utt), ptr(table), ptr(spk_ids), b, l, u, s, ptr(out), k_pad, stream())
    _count()
    return out


def mixer_input_bwd(d_in, spk_ids, batch, frames, s, k_pad, d_table):
    call('srnn_mixer_input_bwd', ptr(d_in), ptr(spk_ids), batch, frames, s, k_pad, ptr(d_table), stream())
    _count()


def tier_input(xq_u8, x_off, lut, frames, conds, batch, t, fs, k_pad):
    l, c = conds.shape[1], conds.shape[2]
    dev = conds.device
    out = torch.empty(batch * t, k_pad, dtype=BF16, device=dev)
    call('srnn_tier_input', ptr(xq_u8), xq_u8.shape[1] if xq_u8 is not None else 0, x_off, ptr(lut), ptr(frames),
         ptr(conds), batch, t, fs, l, c, ptr(out), k_pad, stream())
    _count()
    return out


def tier_input_bwd(d_in, batch, t, fs, l, c, k_pad, dconds):
    call('srnn_tier_input_bwd', ptr(d_in), batch, t, fs, l, c, k_pad, ptr(dconds), stream())
    _count()


def repeat_rows(x, rows, cols, ld_in, rep, out, ld_out):
    call('srnn_repeat_rows', ptr(x), rows, cols, ld_in, rep, ptr(out), ld_out, stream())
    _count()


def repeat_rows_bwd(dout, rows, cols, ld_dout, rep, din, ld_din):
    call('srnn_repeat_rows_bwd', ptr(dout), rows, cols, ld_dout, rep, ptr(din), ld_din, stream())
    _count()


# ----------------------------------------------------------------------------------------------
# GEMMs
# ----------------------------------------------------------------------------------------------
def gemm_nt(a, b, c, m, n, k, lda, ldb, ldc, batch=1, a_bs=0, c_bs=0, bias=None, aux=None, ldaux=0, aux_bs=0,
            aux_mode=0, relu=False, n_fold=0, aux_row_div=1, colsum=None, a2=None, lda2=0, a2_bs=0, k1=0, relu_mask=None,
            gate_mask=None):
    """C_i[m,n] = epi(A_i[m,k] . B[n,k]^T); c.dtype selects bf16 / fp32 output."""
    _need(a, BF16, 'gemm A')
    _need(b, BF16, 'gemm B')
    g = GemmArgs()
    g.op, g.m, g.n, g.k, g.batch = 0, m, n, k, batch
    g.a, g.lda, g.a_batch_stride, g.a_row_offset = a.data_ptr(), lda, a_bs, 0
    g.b, g.ldb, g.b_batch_stride, g.b_row_offset = b.data_ptr(), ldb, 0, 0
    g.c, g.ldc, g.c_batch_stride = c.data_ptr(), ldc, c_bs
    g.c_dtype = 0 if c.dtype == BF16 else 1
    g.n_fold = n_fold
    g.bias = bias.data_ptr() if bias is not None else None
    g.aux = aux.data_ptr() if aux is not None else None
    g.ldaux, g.aux_batch_stride, g.aux_mode = ldaux, aux_bs, aux_mode if aux is not None else 0
    g.relu = int(relu)
    g.aux_row_div = aux_row_div
    g.max_ctas = gemm_max_ctas
    g.colsum = colsum.data_ptr() if colsum is not None else None
    for mk, field in ((relu_mask, 'relu_mask'), (gate_mask, 'gate_mask')):
        if mk is not None:
            _need(mk, torch.int32, 'gemm ' + field)
            setattr(g, field, mk.data_ptr())
            g.ldmask = mk.shape[-1]
    if a2 is not None:
        _need(a2, BF16, 'gemm A2')
        g.a2, g.lda2, g.a2_batch_stride, g.k1 = a2.data_ptr(), lda2, a2_bs, k1
    _lib.profile_note = f'NT m={m}x{batch} n={n} k={k}'
    call('srnn_gemm_bf16', C.byref(g), stream())
    _count()
    return c


def gemm_tn(a, b, c, m, n, k, lda, ldb, ldc, batch=1, a_bs=0, b_bs=0, a_off=0, b_off=0):
    """C[m,n] += sum_i A_i[k,m]^T . B_i[k,n]  (c fp32, must be initialised by the caller)."""
    _need(a, BF16, 'gemm A')
    _need(b, BF16, 'gemm B')
    _need(c, F32, 'gemm C')
    g = GemmArgs()
    g.op, g.m, g.n, g.k, g.batch = 1, m, n, k, batch
    g.a, g.lda, g.a_batch_stride, g.a_row_offset = a.data_ptr(), lda, a_bs, a_off
    g.b, g.ldb, g.b_batch_stride, g.b_row_offset = b.data_ptr(), ldb, b_bs, b_off
    g.c, g.ldc, g.c_batch_stride = c.data_ptr(), ldc, 0
    g.c_dtype = 1
    g.max_ctas = gemm_max_ctas
    _lib.profile_note = f'TN m={m} n={n} k={k}x{batch}'
    call('srnn_gemm_bf16', C.byref(g), stream())
    _count()
    return c


def gemm_nll(mode, a, w, bias, target, m, k, lda, ldw, lse=None, logp_target=None, logp=None, row_grad=None, g=None,
             dlogits=None, ldlogp=256):
    n = NllArgs()
    n.mode, n.m, n.k = mode, m, k
    n.a, n.lda, n.w, n.ldw = a.data_ptr(), lda, w.data_ptr(), ldw
    n.bias = bias.data_ptr() if bias is not None else None
    n.target = target.data_ptr()
    n.lse = lse.data_ptr() if lse is not None else None
    n.logp_target = logp_target.data_ptr() if logp_target is not None else None
    n.logp, n.ldlogp = (logp.data_ptr(), ldlogp) if logp is not None else (None, 0)
    n.row_grad = row_grad.data_ptr() if row_grad is not None else None
    n.g, n.ldg = (g.data_ptr(), 256) if g is not None else (None, 0)
    n.dlogits, n.lddlogits = (dlogits.data_ptr(), 256) if dlogits is not None else (None, 0)
    _lib.profile_note = f'NLL mode={mode} m={m} k={k}'
    call('srnn_gemm_nll', C.byref(n), stream())
    _count()


# ----------------------------------------------------------------------------------------------
# recurrence
# ----------------------------------------------------------------------------------------------
GRU_MAX_BATCH = 64          # rows of one MMA tile (a group)
GRU_MAX_STEP_BATCH = int(os.environ.get('SRNN_GRU_MAX_STEP_BATCH', '512'))    # rows per launch (8 groups); 64: every group its own launch
GRU_SYNC_WORDS = 8192       # srnn_gru_args.sync: arrival counter, statistics and one release-flag line per CTA
gru_tuning_flags = int(os.environ.get('SRNN_GRU_TUNING_FLAGS', '0'))     # srnn_gru_args.tuning_flags (scripts/gru_microbench.py sweeps them)
gru_debug_ts = None      # int64 [256, 8] tensor receiving CTA 0's pipeline timestamps
gru_units_per_cta = 8    # 16 halves the recurrent kernels' CTA count (SMs left free for concurrent GEMMs)


_gru_scratch = {}
gru_last_sync = None
#: callables invoked (host side) right after a recurrent BACKWARD launch has been enqueued.  The data-parallel trainer uses it
#: to enqueue its gradient all-reduces BEHIND that launch: a cooperative launch does not start while another stream's kernel
#: (an NCCL all-reduce waiting for a slower rank) is resident, so an all-reduce enqueued just before it delays the recurrence
#: by however long the slowest rank is late (profiles/r02_dp_trace_n8.txt)
rnn_backward_listeners = []


def _gru_call(name, batch, steps, hidden, cell=0, **bufs):
    """Runs the persistent kernel over the batch rows (independent sequences), at most GRU_MAX_STEP_BATCH per launch.
    ``bufs``: field -> (tensor, elements per batch row); time-major buffers advance by one row."""
    # one launch takes up to GRU_MAX_STEP_BATCH rows: the kernel walks groups of 64 rows inside every timestep (one grid
    # handshake per timestep for all of them, the weight slice fetched once)
    group = GRU_MAX_STEP_BATCH
    for b0 in range(0, batch, group):
        nb = min(group, batch - b0)
        a = GruArgs()
        a.batch, a.steps, a.hidden, a.ext_batch, a.cell = nb, steps, hidden, batch, cell
        for key, (t, per_row) in bufs.items():
            if t is None:
                setattr(a, key, None)
            else:
                setattr(a, key, t.data_ptr() + b0 * per_row * t.element_size())
        dev = bufs['h_ext'][0].device
        if steps == 1 and name == 'srnn_gru_forward':
            # a single FORWARD timestep never waits on the arrival counter: no zeroing needed.  (The backward kernel
            # runs steps + 1 rounds and does wait at round 1, so it always gets a freshly zeroed counter.)
            sync = _gru_scratch.get(dev)
            if sync is None:
                sync = _gru_scratch[dev] = torch.zeros(GRU_SYNC_WORDS, dtype=torch.int32, device=dev)
        else:
            sync = torch.zeros(GRU_SYNC_WORDS, dtype=torch.int32, device=dev)
        a.sync = sync.data_ptr()
        global gru_last_sync
        gru_last_sync = sync               # [32] = exchange attempts the launch rejected and repeated (tests read it)
        a.tuning_flags = gru_tuning_flags
        a.units_per_cta = gru_units_per_cta
        a.debug_ts = gru_debug_ts.data_ptr() if gru_debug_ts is not None else None
        _lib.profile_note = f'B={nb} T={steps} H={hidden}' + (' lstm' if cell else '')
        with timed(f"{'rnn_fwd' if name == 'srnn_gru_forward' else 'rnn_bwd'}_T{steps}"):
            call(name, C.byref(a), stream())
        _count(1 if steps == 1 else 2)
    if name == 'srnn_gru_backward':
        for fn in rnn_backward_listeners:
            fn()


def gru_forward(gi, w_hh, b_hh, h_ext, hall, h_state, gates, batch, steps, hidden):
    """h_ext: bf16 [steps+1, batch, H] time-major (slot 0 = initial state); hall: bf16 [batch*steps, H]."""
    h = hidden
    _gru_call('srnn_gru_forward', batch, steps, h, gi=(gi, steps * 3 * h), w_hh=(w_hh, 0), b_hh=(b_hh, 0),
              h_ext=(h_ext, h), hall=(hall, steps * h), h_state=(h_state, h), gates=(gates, steps * 4 * h))


def gru_backward(w_hh_t, h_ext, gates, dh_out, dgi, dgh, dh0, batch, steps, hidden, db_ih=None, db_hh=None):
    """dgh: bf16 [steps, batch, 3H] time-major; dgi: bf16 [batch*steps, 3H] batch-major; db_ih / db_hh: optional fp32
    [3H] bias gradients, ACCUMULATED into (the caller zeroes them)."""
    h = hidden
    _gru_call('srnn_gru_backward', batch, steps, h, w_hh=(w_hh_t, 0), h_ext=(h_ext, h),
              gates=(gates, steps * 4 * h), dh_out=(dh_out, steps * h), dgi=(dgi, steps * 3 * h),
              dgh=(dgh, 3 * h), dh0=(dh0, h), db_ih=(db_ih, 0), db_hh=(db_hh, 0))


def lstm_forward(gi, w_hh, b_hh, h_ext, hall, h_state, c_state, gates, batch, steps, hidden):
    """LSTM extension: gi [batch*steps, 4H], w_hh [4H, H], gates [batch*steps, 5H] (i,f,g,o,c)."""
    h = hidden
    _gru_call('srnn_gru_forward', batch, steps, h, cell=1, gi=(gi, steps * 4 * h), w_hh=(w_hh, 0), b_hh=(b_hh, 0),
              h_ext=(h_ext, h), hall=(hall, steps * h), h_state=(h_state, h), c_state=(c_state, h),
              gates=(gates, steps * 5 * h))


def lstm_backward(w_hh_t, h_ext, gates, c_init, dh_out, dgi, dgh, dh0, dc0, batch, steps, hidden, db_ih=None, db_hh=None):
    """w_hh_t [H, 4H]; dgh [steps, batch, 4H] time-major; dgi [batch*steps, 4H] batch-major."""
    h = hidden
    _gru_call('srnn_gru_backward', batch, steps, h, cell=1, w_hh=(w_hh_t, 0), h_ext=(h_ext, h),
              gates=(gates, steps * 5 * h), c_init=(c_init, h), dh_out=(dh_out, steps * h),
              dgi=(dgi, steps * 4 * h), dgh=(dgh, 4 * h), dh0=(dh0, h), dc0=(dc0, h), db_ih=(db_ih, 0), db_hh=(db_hh, 0))


def state_select(carried, h0, use_carry, batch, hidden):
    h_state = torch.empty(batch, hidden, dtype=F32, device=h0.device)
    call('srnn_state_select', ptr(carried), ptr(h0), ptr(use_carry), batch, hidden, ptr(h_state), None, 0, stream())
    _count()
    return h_state


def state_select_bwd(dh, use_carry, batch, hidden):
    out = torch.empty(hidden, dtype=F32, device=dh.device)
    call('srnn_state_select_bwd', ptr(dh), ptr(use_carry), batch, hidden, ptr(out), stream())
    _count()
    return out


def masked_nll_mean(logp_target, slot_valid, rows_per_slot):
    out = torch.empty(2, dtype=F32, device=logp_target.device)
    call('srnn_masked_nll_mean', ptr(logp_target), ptr(slot_valid), logp_target.numel(), rows_per_slot, ptr(out), stream())
    _count(3)
    return out


def adam_clipped(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0):
    call('srnn_adam_clipped', ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), float(lr),
         float(beta1), float(beta2), float(eps), int(step), float(grad_scale), stream())
    _count()


# ----------------------------------------------------------------------------------------------
# fp32-tolerance mode (csrc/precise.cu): fp32 tensors everywhere, split-bf16 operands into the tcgen05 GEMM
# ----------------------------------------------------------------------------------------------
def _mat(x, name):
    """2-D fp32 CUDA view with unit column stride -> (tensor, rows, cols, ld)."""
    _need(x, F32, name)
    if x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError(f'{name}: expected a 2-D matrix with unit column stride, got {tuple(x.shape)} / {x.stride()}')
    return x, x.shape[0], x.shape[1], x.stride(0)


SPLIT_SEGMENTS = {0: 3, 1: 3, 2: 2, 3: 6, 4: 6}


def split3(x, role):
    """fp32 (rows, cols) -> bf16 (rows, n_seg * cols_pad): role 0 [lo|hi|hi], 1 [hi|lo|hi], 2 [lo|hi] (two pieces, 2^-18);
    role 3 [p2|p1|p0|p1|p0|p0], 4 [q0|q1|q2|q0|q1|q0] (three pieces, fp32-exact operands, six products) - the smallest
    products come first along K because the tensor core's accumulator truncates; cols_pad = cols rounded up to 8.
    Returns (tensor, cols_pad)."""
    x, rows, cols, ld = _mat(x, 'split3 x')
    kp = round_up(cols, 8)
    seg = SPLIT_SEGMENTS[role]
    out = torch.empty(rows, seg * kp, dtype=BF16, device=x.device)
    call('srnn_split3_bf16', ptr(x), rows, cols, ld, ptr(out), kp, seg * kp, role, stream())
    _count()
    return out, kp


def gemm_nt32(a, w, out=None, bias=None, gate_mask=None, colsum=None, terms=3):
    """out[m,n] = a[m,k] . w[n,k]^T (+ bias) at fp32-level accuracy: ONE bf16 GEMM over K' = terms * k on split operands.
    terms = 3: two pieces per operand, products accurate to 4.5e-6 (the 2^-18 of the split); terms = 6: three pieces,
    8e-8 at short K, 1e-6 at K = 1024 (what is left is the tensor core's truncating fp32 accumulation, ~3e-8 per K=16
    update of the LAST segment - the segments are ordered smallest product first).  The model path uses 6 forward (ReLU
    gates flip on pre-activation errors) and 3 backward.
    ``a``/``w``/``out`` are fp32 matrices (views with a row stride are fine)."""
    a, m, k, _ = _mat(a, 'gemm_nt32 a')
    w, n, k2, _ = _mat(w, 'gemm_nt32 w')
    if k != k2:
        raise RuntimeError(f'gemm_nt32: inner dimensions differ ({k} vs {k2})')
    if out is None:
        out = torch.empty(m, n, dtype=F32, device=a.device)
    out, _, _, ldc = _mat(out, 'gemm_nt32 out')
    a3, kp = split3(a, 0 if terms == 3 else 3)
    w3, _ = split3(w, 1 if terms == 3 else 4)
    gemm_nt(a3, w3, out, m, n, terms * kp, terms * kp, terms * kp, ldc, bias=bias, gate_mask=gate_mask, colsum=colsum)
    return out


def embed_gather_f32(table, idx_u8, rows_per_slot, r0, q):
    """out[(b,j), :] = sum_k table[k*q + idx[b, j+k], :]; table fp32 (r0*q, H) contiguous, idx uint8 (B, W)."""
    _need(table, F32, 'embed_gather_f32 table')
    _need(idx_u8, torch.uint8, 'embed_gather_f32 idx')
    b, w = idx_u8.shape
    h = table.shape[1]
    out = torch.empty(b * rows_per_slot, h, dtype=F32, device=table.device)
    call('srnn_embed_gather_f32', ptr(table), ptr(idx_u8), w, b, rows_per_slot, r0, q, h, ptr(out), stream())
    _count()
    return out


def gemm_tn32(a, b, out):
    """out[m,n] += a[rows,m]^T . b[rows,n] at fp32-level accuracy (three accumulating TN GEMMs on the split segments)."""
    a, rows, m, _ = _mat(a, 'gemm_tn32 a')
    b, rows2, n, _ = _mat(b, 'gemm_tn32 b')
    if rows != rows2:
        raise RuntimeError(f'gemm_tn32: contraction lengths differ ({rows} vs {rows2})')
    out, _, _, ldc = _mat(out, 'gemm_tn32 out')
    a3, mp = split3(a, 0)
    b3, np_ = split3(b, 1)
    for i in range(3):
        gemm_tn(a3[:, i * mp:], b3[:, i * np_:], out, m, n, rows, 3 * mp, 3 * np_, ldc)
    return out


def mixer_input_f32(utt, table, spk_ids, k_pad):
    b, l, u = utt.shape
    out = torch.empty(b * l, k_pad, dtype=F32, device=utt.device)
    call('srnn_mixer_input_f32', ptr(utt), ptr(table), ptr(spk_ids), b, l, u, table.shape[1], ptr(out), k_pad, stream())
    _count()
    return out


def mixer_input_bwd_f32(d_in, spk_ids, batch, frames, s, k_pad, d_table):
    _need(d_in, F32, 'mixer_input_bwd_f32 d_in')
    call('srnn_mixer_input_bwd_f32', ptr(d_in), ptr(spk_ids), batch, frames, s, k_pad, ptr(d_table), stream())
    _count()


def tier_input_f32(xq_u8, x_off, lut, frames, conds, batch, t, fs, k_pad):
    l, c = conds.shape[1], conds.shape[2]
    out = torch.empty(batch * t, k_pad, dtype=F32, device=conds.device)
    call('srnn_tier_input_f32', ptr(xq_u8), xq_u8.shape[1] if xq_u8 is not None else 0, x_off, ptr(lut), ptr(frames),
         ptr(conds), batch, t, fs, l, c, ptr(out), k_pad, stream())
    _count()
    return out


def tier_input_bwd_f32(d_in, batch, t, fs, l, c, dconds):
    d_in, _, _, ld = _mat(d_in, 'tier_input_bwd_f32 d_in')
    call('srnn_tier_input_bwd_f32', ptr(d_in), batch, t, fs, l, c, ld, ptr(dconds), stream())
    _count()


def weight_prep_f32(v, g, shape3, out1, s1, out2=None, s2=None, inv_norm=None):
    """``weight_prep`` with fp32 outputs."""
    _need(v, F32, 'weight_prep_f32 v')
    _need(out1, F32, 'weight_prep_f32 out1')
    r, a, b = shape3
    assert v.is_contiguous() and v.numel() == r * a * b
    call('srnn_weight_prep_f32', ptr(v), ptr(g), r, a, b, ptr(out1), _strides(s1), ptr(out2),
         _strides(s2) if s2 is not None else None, ptr(inv_norm), stream())
    _count()


def bias_act_f32(x, aux=None, aux_row_div=1, aux2=None, relu=False, mask=None):
    """in place: x = act(x + aux[row // aux_row_div] + aux2); ``mask`` (int32 (rows, ceil(cols/32))) receives result > 0."""
    x, rows, cols, ld = _mat(x, 'bias_act_f32 x')
    ldaux = ldaux2 = 0
    if aux is not None:
        aux, _, _, ldaux = _mat(aux, 'bias_act_f32 aux')
    if aux2 is not None:
        aux2, _, _, ldaux2 = _mat(aux2, 'bias_act_f32 aux2')
    if mask is not None:
        _need(mask, torch.int32, 'bias_act_f32 mask')
    call('srnn_bias_act_f32', ptr(x), rows, cols, ld, ptr(aux), ldaux, aux_row_div, ptr(aux2), ldaux2, int(relu),
         ptr(mask), mask.shape[-1] if mask is not None else 0, stream())
    _count()
    return x


def segment_sum_f32(dout, rep):
    dout, rows_out, cols, ld = _mat(dout, 'segment_sum_f32 dout')
    rows = rows_out // rep
    din = torch.empty(rows, cols, dtype=F32, device=dout.device)
    call('srnn_segment_sum_f32', ptr(dout), rows, cols, ld, rep, ptr(din), cols, stream())
    _count()
    return din


def colsum_f32(x):
    x, rows, cols, ld = _mat(x, 'colsum_f32 x')
    out = torch.empty(cols, dtype=F32, device=x.device)
    call('srnn_colsum_f32', ptr(x), rows, cols, ld, ptr(out), stream())
    _count(2)
    return out


def logsoftmax_nll_f32(logits, target_u8):
    """in place: logits -> log-probabilities; returns (lse, logp_target)."""
    logits, m, q, ld = _mat(logits, 'logsoftmax_nll_f32 logits')
    lse = torch.empty(m, dtype=F32, device=logits.device)
    logp_t = torch.empty(m, dtype=F32, device=logits.device)
    call('srnn_logsoftmax_nll_f32', ptr(logits), ld, m, q, ptr(target_u8), ptr(lse), ptr(logp_t), stream())
    _count()
    return lse, logp_t


def logsoftmax_nll_bwd_f32(logp, target_u8, row_grad=None, g=None):
    logp, m, q, ld = _mat(logp, 'logsoftmax_nll_bwd_f32 logp')
    dl = torch.empty(m, q, dtype=F32, device=logp.device)
    ldg = 0
    if g is not None:
        g, _, _, ldg = _mat(g, 'logsoftmax_nll_bwd_f32 g')
    call('srnn_logsoftmax_nll_bwd_f32', ptr(logp), ld, m, q, ptr(target_u8), ptr(row_grad), ptr(g), ldg, ptr(dl), q,
         stream())
    _count()
    return dl


def gru_forward_f32(gi, w_hh, b_hh, h_state, batch, steps, hidden):
    """gi fp32 [batch*steps, 3H]; w_hh fp32 [3H, H]; h_state fp32 [batch, H] (in: initial, out: final).
    Returns (hall [batch*steps, H], gates [batch*steps, 4H])."""
    h = hidden
    dev = gi.device
    w3, _ = split3(w_hh, 4)                                  # forward: three pieces, six products (see gemm_nt32)
    a = _lib.GruF32Args()
    a.batch, a.steps, a.hidden = batch, steps, h
    hall = torch.empty(batch * steps, h, dtype=F32, device=dev)
    gates = torch.empty(batch * steps, 4 * h, dtype=F32, device=dev)
    a3 = torch.empty(batch, 6 * h, dtype=BF16, device=dev)
    ws = torch.empty(batch, 3 * h, dtype=F32, device=dev)
    a.gi, a.w3, a.b_hh, a.h_state = gi.data_ptr(), w3.data_ptr(), b_hh.data_ptr(), h_state.data_ptr()
    a.hall, a.gates, a.a3, a.ws = hall.data_ptr(), gates.data_ptr(), a3.data_ptr(), ws.data_ptr()
    with timed(f'rnn_fwd_f32_T{steps}'):
        call('srnn_gru_forward_f32', C.byref(a), stream())
    _count(1 + 2 * steps)
    return hall, gates


def gru_backward_f32(w_hh, gates, hall, h_init, dh_out, batch, steps, hidden):
    """-> (dgi, dgh) fp32 [batch*steps, 3H] batch-major, dh0 fp32 [batch, H]."""
    h = hidden
    dev = gates.device
    w3, _ = split3(w_hh.t().contiguous(), 1)                 # W_hh^T [H, 3H] -> bf16 [H, 9H]
    a = _lib.GruF32Args()
    a.batch, a.steps, a.hidden = batch, steps, h
    dgi = torch.empty(batch * steps, 3 * h, dtype=F32, device=dev)
    dgh = torch.empty(batch * steps, 3 * h, dtype=F32, device=dev)
    dh0 = torch.empty(batch, h, dtype=F32, device=dev)
    carry = torch.empty(batch, h, dtype=F32, device=dev)
    a3 = torch.empty(batch, 9 * h, dtype=BF16, device=dev)
    ws = torch.empty(batch, h, dtype=F32, device=dev)
    a.w3, a.hall, a.h_init, a.gates = w3.data_ptr(), hall.data_ptr(), h_init.data_ptr(), gates.data_ptr()
    a.a3, a.ws, a.dh_out, a.dgi, a.dgh = a3.data_ptr(), ws.data_ptr(), dh_out.data_ptr(), dgi.data_ptr(), dgh.data_ptr()
    a.dh0, a.carry = dh0.data_ptr(), carry.data_ptr()
    with timed(f'rnn_bwd_f32_T{steps}'):
        call('srnn_gru_backward_f32', C.byref(a), stream())
    _count(3 + 2 * steps)
    return dgi, dgh, dh0
