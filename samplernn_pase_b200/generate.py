"""Autoregressive generation (``SampleRNNModel.test``, model.py:289-351; BASELINE config 5).

Batched, sample-by-sample generation on the same C-ABI kernels as training: per output sample the
frame tiers whose frame boundary is reached run one recurrent step (persistent GRU kernel with
steps=1) and the sample-level MLP runs on the last r0 generated samples.  All weights are prepared
(weight-norm, bf16 GEMM layouts, the one-hot x embedding table) ONCE per call, not per sample.

This is the first correct path for generation; a persistent per-sample kernel that
keeps the weights on chip is the performance follow-up.
"""
import torch

from . import ops
from .ops import BF16, F32, round_up


def _e(*shape, dtype=BF16, device=None):
    return torch.empty(*shape, dtype=dtype, device=device)


def _z(*shape, dtype=BF16, device=None):
    return torch.zeros(*shape, dtype=dtype, device=device)


class TierWeights:
    """bf16 GEMM-layout weights of one FrameLevelLayer (model.py:96-156)."""

    def __init__(self, layer, conds_size):
        dev = layer.rnn_h0.device
        h, fs, r, c = layer.rnn_hidden_size, layer.input_samples, layer.ratio, conds_size
        self.h, self.fs, self.r, self.layers = h, fs, r, layer.rnn_layers
        self.kp = round_up(fs + c, 8)
        self.wcat = _z(h, self.kp, device=dev)
        ops.weight_prep(layer.x_expand.weight_v.detach(), layer.x_expand.weight_g.detach(), (h, fs, 1), self.wcat,
                        (self.kp, 1, 0))
        ops.weight_prep(layer.conds_expand.weight_v.detach(), layer.conds_expand.weight_g.detach(), (h, c, 1),
                        self.wcat[:, fs:], (self.kp, 1, 0))
        self.bias_u = (layer.x_expand.bias + layer.conds_expand.bias).detach().contiguous()
        self.rnn = []
        self.lstm = layer.rnn_cell == 'lstm'
        ng = 4 if self.lstm else 3
        self.ng = ng
        flat = layer.rnn.flat(layer.rnn_layers)
        for i in range(layer.rnn_layers):
            w_ih, w_hh, b_ih, b_hh = (t.detach().contiguous() for t in flat[4 * i: 4 * i + 4])
            wih, whh = _e(ng * h, h, device=dev), _e(ng * h, h, device=dev)
            ops.weight_prep(w_ih, None, (ng * h, h, 1), wih, (h, 1, 0))
            ops.weight_prep(w_hh, None, (ng * h, h, 1), whh, (h, 1, 0))
            self.rnn.append((wih, whh, b_ih, b_hh))
        self.wu = _e(r * h, h, device=dev)
        ops.weight_prep(layer.upsample.weight_v.detach(), layer.upsample.weight_g.detach(), (h, h, r), self.wu,
                        (1, h, h * h))
        self.bias_up = layer.upsample_bias.detach().t().contiguous().view(-1)
        self.h0 = layer.rnn_h0.detach()
        self.c0 = layer.rnn_c0.detach() if self.lstm else None


class SampleWeights:
    """bf16 GEMM-layout weights of the SampleLevelLayer (model.py:159-203)."""

    def __init__(self, sl, conds_size):
        dev = sl.emb_layer.weight.device
        ev = sl.emb_layer_expand.weight_v.detach()
        h, q, r0 = ev.shape
        c = conds_size
        self.h, self.q, self.r0, self.cp = h, q, r0, round_up(c, 8)
        e_b = ops.to_bf16(sl.emb_layer.weight.detach())
        we = _e(h, r0 * q, device=dev)
        ops.weight_prep(ev, sl.emb_layer_expand.weight_g.detach(), (h, q, r0), we, (r0 * q, 1, q))
        self.table = _e(h, r0 * q, device=dev)                     # T[o, k*Q+q] = sum_q' We[o,q',k] E[q,q']
        for k in range(r0):
            ops.gemm_nt(we[:, k * q:], e_b, self.table[:, k * q:], h, q, q, r0 * q, q, r0 * q)
        self.wcomb = _e(h, 3 * h, device=dev)
        ops.weight_prep(sl.comb_layer.weight.detach().contiguous(), None, (h, 3 * h, 1), self.wcomb, (3 * h, 1, 0))
        # folded tables for one-step generation: tt[k*Q+q, o] = sum_q' E[q,q'] We[o,q',k] (embedding + conv tap k of
        # code q), then through the embedding block of comb_layer: table_t[k*Q+q, :] = Wcomb[:, :H] . tt[k*Q+q, :]
        tt = _e(r0 * q, h, device=dev)
        for k in range(r0):
            ops.gemm_nt(e_b, we[:, k * q:], tt[k * q:], q, h, q, q, r0 * q, h)
        self.table_t = _e(r0 * q, h, device=dev)
        ops.gemm_nt(tt, self.wcomb, self.table_t, r0 * q, h, h, h, 3 * h, h)
        self.wcs = _z(h, self.cp, device=dev)
        ops.weight_prep(sl.conds_expand.weight.detach().contiguous(), None, (h, c, 1), self.wcs, (self.cp, 1, 0))
        self.csb = sl.conds_expand.bias.detach().contiguous()
        self.cbias = sl.comb_layer.bias.detach().contiguous()
        self.w2 = _e(h, h, device=dev)
        ops.weight_prep(sl.comb_layer_expand.weight_v.detach(), sl.comb_layer_expand.weight_g.detach(), (h, h, 1),
                        self.w2, (h, 1, 0))
        self.b2 = sl.comb_layer_expand.bias.detach().contiguous()
        self.w3 = _e(q, h, device=dev)
        ops.weight_prep(sl.adapt.weight_v.detach(), sl.adapt.weight_g.detach(), (q, h, 1), self.w3, (h, 1, 0))
        self.b3 = sl.adapt.bias.detach().contiguous()


def tier_step(w, lut, win, conds_row, upper, upper_ld, h_state, out, c_state=None):
    """One frame of a tier for every utterance.  ``win`` (B, FS) uint8 holds the last FS generated samples (the
    tier reads its last fs), ``conds_row`` (B,1,C) fp32, ``upper`` a (B,H) bf16 view with row stride ``upper_ld``
    (or None for the top tier), ``h_state`` (and for LSTM tiers ``c_state``) (layers,B,H) fp32 updated in place.
    Writes the r upsampled conditioning vectors into ``out`` (B, r, H) bf16."""
    b, fs_top = win.shape
    dev = win.device
    h = w.h
    ain = ops.tier_input(win, fs_top - w.fs, lut, None, conds_row, b, 1, w.fs, w.kp)
    x = _e(b, h, device=dev)
    ops.gemm_nt(ain, w.wcat, x, b, h, w.kp, w.kp, w.kp, h, bias=w.bias_u, aux=upper, ldaux=upper_ld, aux_mode=1)
    for i, (wih, whh, b_ih, b_hh) in enumerate(w.rnn):
        gi = _e(b, w.ng * h, device=dev)
        ops.gemm_nt(x, wih, gi, b, w.ng * h, h, h, h, w.ng * h, bias=b_ih)
        h_ext = _e(2, b, h, device=dev)
        ops.pad_cast_bf16(h_state[i], b, h, h, h_ext, h, h)
        hall = _e(b, h, device=dev)
        if w.lstm:
            gates = _e(b, 5 * h, device=dev)
            ops.lstm_forward(gi, whh, b_hh, h_ext, hall, h_state[i], c_state[i], gates, b, 1, h)
        else:
            gates = _e(b, 4 * h, device=dev)
            ops.gru_forward(gi, whh, b_hh, h_ext, hall, h_state[i], gates, b, 1, h)
        x = hall
    ops.gemm_nt(x, w.wu, out, b, w.r * h, h, h, h, w.r * h, bias=w.bias_up)
    return out


def frame_terms(sw, conds_b, c_term, cc):
    """Per top-tier frame: ``c_term`` = conds_expand(conds) (model.py:194) and ``cc`` = the conditioning block of
    comb_layer applied to it, plus comb_layer's bias (model.py:195-200) - both constant over the frame."""
    b, h, cp = conds_b.shape[0], sw.h, sw.cp
    ops.gemm_nt(conds_b, sw.wcs, c_term, b, h, cp, cp, cp, h, bias=sw.csb)
    ops.gemm_nt(c_term, sw.wcomb[:, h:], cc, b, h, h, h, 3 * h, h, bias=sw.cbias)


def sample_pre(sw, up0, cc, pre):
    """Per lowest-tier frame: pre[b, j] = comb_layer's upper-tier block applied to the j-th upsampled vector + cc[b]
    (everything of comb_layer's input that does not depend on the samples of the frame)."""
    b, r0, h = up0.shape
    ops.gemm_nt(up0, sw.wcomb[:, 2 * h:], pre, b * r0, h, h, h, 3 * h, h, aux=cc, ldaux=h, aux_mode=1, aux_row_div=r0)


def sample_step(sw, win, pre_j, pre_ld, logits):
    """Logits (B,256) fp32 of the next sample: the embedding side is a sum of r0 table rows selected by the last r0
    samples of ``win`` (srnn_embed_sum), then comb_layer_expand and adapt (model.py:201-202); the log-softmax
    (model.py:203) is taken by the sampling kernel."""
    b, fs_top = win.shape
    dev = win.device
    h, q, r0 = sw.h, sw.q, sw.r0
    h1 = _e(b, h, device=dev)
    ops.embed_sum(sw.table_t, win[:, fs_top - r0:], fs_top, b, r0, q, h, pre_j, pre_ld, True, h1, h)
    h2 = _e(b, h, device=dev)
    ops.gemm_nt(h1, sw.w2, h2, b, h, h, h, h, h, bias=sw.b2, relu=True)
    ops.gemm_nt(h2, sw.w3, logits, b, q, h, h, h, q, bias=sw.b3)
    return logits


#: set True to decode greedily (argmax) instead of sampling - lets eager and CUDA-graph runs be compared exactly
_GREEDY = False


class _GenState:
    """Static buffers of a generation call (CUDA-graph capturable step programs run against these)."""

    def __init__(self, b, fs_top, tiers, sw, c, return_logp, dev):
        self.win = None
        self.conds_cur = torch.empty(b, 1, c, dtype=F32, device=dev)
        self.conds_b = _e(b, sw.cp, device=dev)
        self.c_term = _e(b, sw.h, device=dev)
        self.cc = _e(b, sw.h, device=dev)
        self.outs = [_e(b, w.r, w.h, device=dev) for w in tiers]
        self.pre = _e(b, sw.r0, sw.h, device=dev)
        self.frame_out = torch.empty(b, fs_top, dtype=torch.uint8, device=dev)
        self.logp_frame = torch.empty(b, fs_top, sw.q, dtype=F32, device=dev) if return_logp else None
        self.logits = torch.empty(b, sw.q, dtype=F32, device=dev)
        self.rng_state = torch.zeros(3, dtype=torch.int64, device=dev)      # {seed, step, 0} of the device-side draws


def _frame_phase(p, tiers, sw, lut, st, states, cstates):
    """One sample step at phase ``p = xi % FS`` of a top-tier frame, written against the static buffers of ``st`` so
    that it can be captured in a CUDA graph: ``st.win`` (B,FS) holds the last FS generated samples, ``st.outs`` the
    tiers' current upsampled outputs, ``st.frame_out[:, p]`` receives the new sample."""
    win = st.win
    b, fs_top = win.shape
    for n in reversed(range(len(tiers))):                                    # model.py:312-337
        w = tiers[n]
        if p % w.fs != 0:
            continue
        upper, upper_ld = None, 0
        if n != len(tiers) - 1:
            frame_index = (p % tiers[n + 1].fs) // w.fs                      # == (xi // fs_n) % r_{n+1}
            upper, upper_ld = st.outs[n + 1][:, frame_index], tiers[n + 1].r * w.h
        tier_step(w, lut, win, st.conds_cur, upper, upper_ld, states[n], st.outs[n], cstates[n])
        if n == 0:
            sample_pre(sw, st.outs[0], st.cc, st.pre)
    j = p % sw.r0                                                            # model.py:343
    sample_step(sw, win, st.pre[:, j], sw.r0 * sw.h, st.logits)
    # model.py:203,346-348: log-softmax, draw from it, append to the window of the last FS samples
    ops.sample_categorical(st.logits, b, sw.q, None, win, fs_top, st.frame_out[:, p], fs_top, normalise=True,
                           logp_out=st.logp_frame[:, p] if st.logp_frame is not None else None,
                           rng_state=None if _GREEDY else st.rng_state)


@torch.no_grad()
def generate(model, utt_conds, info, return_logp=False, generator=None, use_graphs=True):
    """model.py:289-351, batched over the first dimension of ``utt_conds`` (the reference handles one
    utterance per call).  Returns int64 (B, (t+1)*FS) with FS leading ``quantize_zero()`` samples; with
    ``return_logp`` also the (B, t*FS, Q) log-probabilities each sample was drawn from.

    The FS sample steps of one top-tier frame always run the same kernels on the same buffers (which tier
    fires and which upsampled vector is read depends only on ``xi % FS``), so after an eager first frame the
    FS step programs are captured once as ONE CUDA graph and replayed for every later frame (one replay per FS
    samples keeps the host out of the way: a replay per sample was host-bound at ~40 us per step).  The draws are
    Philox numbers generated inside the sampling kernel (seeded once per call from ``generator``), so the captured
    step programs contain the whole sample step, RNG included."""
    dev = utt_conds.device
    b, t, _ = utt_conds.shape
    infos = info if isinstance(info, (list, tuple)) else [info] * b
    conds = model.conds_mixer(utt_conds, infos).contiguous()                 # (B, t, C) fp32
    c = conds.shape[2]
    fs_top = int(model.frame_size)
    q = model.quantizer
    lut = q.lut(dev)
    total = (t + 1) * fs_top
    y = torch.full((b, total), q.quantize_zero(), dtype=torch.uint8, device=dev)
    tiers = [TierWeights(layer, c) for layer in model.frames_layers]
    sw = SampleWeights(model.sample_layer, c)
    # learnable h0 (model.py:111); clone(): for b == 1 expand().contiguous() would alias the parameter itself
    states = [w.h0[:, None, :].expand(-1, b, -1).clone() for w in tiers]
    cstates = [w.c0[:, None, :].expand(-1, b, -1).clone() if w.lstm else None for w in tiers]   # LSTM extension
    st = _GenState(b, fs_top, tiers, sw, c, return_logp, dev)
    st.win = y[:, :fs_top].clone()                                           # the FS samples before the frame
    # seed of the device-side Philox draws, taken once from ``generator`` (or torch's default CUDA generator)
    if generator is not None and generator.device.type == 'cpu':
        st.rng_state[0:1].copy_(torch.empty(1, dtype=torch.int64).random_(0, 2 ** 62, generator=generator))
    else:
        st.rng_state[0:1].random_(0, 2 ** 62, generator=generator)
    graphed = use_graphs and t > 2
    ops.set_pdl(_PDL)
    try:
        return _generate_frames(model, st, states, cstates, tiers, sw, lut, conds, y, t, b, c, fs_top, graphed, generator,
                                return_logp)
    finally:
        ops.set_pdl(False)


#: launch the per-sample kernels with programmatic dependent launch (their launch / set-up overlaps the predecessor)
_PDL = True


def _generate_frames(model, st, states, cstates, tiers, sw, lut, conds, y, t, b, c, fs_top, graphed, generator, return_logp):
    logps = [] if return_logp else None
    graphs = None
    for f in range(t):                                                       # top-tier frames; xi = (f+1)*FS + p
        st.conds_cur.copy_(conds[:, f: f + 1])                               # model.py:308-309: conds index xi//FS - 1
        ops.pad_cast_bf16(st.conds_cur.view(b, c), b, c, c, st.conds_b, sw.cp, sw.cp)
        frame_terms(sw, st.conds_b, st.c_term, st.cc)
        if graphed and f == 1:                                               # frame 0 ran eagerly (lazy init done)
            torch.cuda.synchronize()
            graphs = torch.cuda.CUDAGraph()                                  # ONE graph = the FS step programs of a frame
            before = ops.launch_count
            with torch.cuda.graph(graphs):
                for p in range(fs_top):
                    _frame_phase(p, tiers, sw, lut, st, states, cstates)
            per_frame = ops.launch_count - before                            # kernels inside the captured graph
            ops.launch_count = before
        if graphs is not None:
            graphs.replay()
            ops.launch_count += per_frame
        else:
            for p in range(fs_top):
                _frame_phase(p, tiers, sw, lut, st, states, cstates)
        y[:, (f + 1) * fs_top: (f + 2) * fs_top] = st.frame_out
        if return_logp:
            logps.append(st.logp_frame.clone())
    out = y.to(torch.int64)
    if return_logp:
        return out, torch.cat(logps, dim=1)
    return out
