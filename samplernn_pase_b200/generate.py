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
        flat = layer.rnn.flat(layer.rnn_layers)
        for i in range(layer.rnn_layers):
            w_ih, w_hh, b_ih, b_hh = (t.detach().contiguous() for t in flat[4 * i: 4 * i + 4])
            wih, whh = _e(3 * h, h, device=dev), _e(3 * h, h, device=dev)
            ops.weight_prep(w_ih, None, (3 * h, h, 1), wih, (h, 1, 0))
            ops.weight_prep(w_hh, None, (3 * h, h, 1), whh, (h, 1, 0))
            self.rnn.append((wih, whh, b_ih, b_hh))
        self.wu = _e(r * h, h, device=dev)
        ops.weight_prep(layer.upsample.weight_v.detach(), layer.upsample.weight_g.detach(), (h, h, r), self.wu,
                        (1, h, h * h))
        self.bias_up = layer.upsample_bias.detach().t().contiguous().view(-1)
        self.h0 = layer.rnn_h0.detach()


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
        self.wcs = _z(h, self.cp, device=dev)
        ops.weight_prep(sl.conds_expand.weight.detach().contiguous(), None, (h, c, 1), self.wcs, (self.cp, 1, 0))
        self.csb = sl.conds_expand.bias.detach().contiguous()
        self.wcomb = _e(h, 3 * h, device=dev)
        ops.weight_prep(sl.comb_layer.weight.detach().contiguous(), None, (h, 3 * h, 1), self.wcomb, (3 * h, 1, 0))
        self.cbias = sl.comb_layer.bias.detach().contiguous()
        self.w2 = _e(h, h, device=dev)
        ops.weight_prep(sl.comb_layer_expand.weight_v.detach(), sl.comb_layer_expand.weight_g.detach(), (h, h, 1),
                        self.w2, (h, 1, 0))
        self.b2 = sl.comb_layer_expand.bias.detach().contiguous()
        self.w3 = _e(q, h, device=dev)
        ops.weight_prep(sl.adapt.weight_v.detach(), sl.adapt.weight_g.detach(), (q, h, 1), self.w3, (h, 1, 0))
        self.b3 = sl.adapt.bias.detach().contiguous()


def tier_step(w, lut, prev_u8, conds_row, upper, h_state):
    """One frame of a tier for every utterance: ``prev_u8`` (B, fs) the fs samples before the frame,
    ``conds_row`` (B,1,C) fp32, ``upper`` (B,H) bf16 or None, ``h_state`` (layers,B,H) fp32 updated in place.
    Returns the r upsampled conditioning vectors (B, r, H) bf16."""
    b = prev_u8.shape[0]
    dev = prev_u8.device
    h = w.h
    ain = ops.tier_input(prev_u8, 0, lut, None, conds_row, b, 1, w.fs, w.kp)
    x = _e(b, h, device=dev)
    ops.gemm_nt(ain, w.wcat, x, b, h, w.kp, w.kp, w.kp, h, bias=w.bias_u, aux=upper, ldaux=h, aux_mode=1)
    for i, (wih, whh, b_ih, b_hh) in enumerate(w.rnn):
        gi = _e(b, 3 * h, device=dev)
        ops.gemm_nt(x, wih, gi, b, 3 * h, h, h, h, 3 * h, bias=b_ih)
        h_ext = _e(2, b, h, device=dev)
        ops.pad_cast_bf16(h_state[i], b, h, h, h_ext, h, h)
        hall = _e(b, h, device=dev)
        gates = _e(b, 4 * h, device=dev)
        ops.gru_forward(gi, whh, b_hh, h_ext, hall, h_state[i], gates, b, 1, h)
        x = hall
    up = _e(b, w.r, h, device=dev)
    ops.gemm_nt(x, w.wu, up, b, w.r * h, h, h, h, w.r * h, bias=w.bias_up)
    return up


def sample_step(w, last_u8, c_term, upper):
    """log-probabilities (B,256) of the next sample given the last r0 samples (B,r0) uint8, the
    conditioning term ``c_term`` (B,H) bf16 (conds_expand of the current frame) and ``upper`` (B,H)."""
    b = last_u8.shape[0]
    dev = last_u8.device
    h, q, r0 = w.h, w.q, w.r0
    onehot = ops.onehot_rows(last_u8.contiguous(), q)              # (B, r0, Q) == (B, r0*Q)
    cat = _e(b, 3 * h, device=dev)
    ops.gemm_nt(onehot, w.table, cat, b, h, r0 * q, r0 * q, r0 * q, 3 * h)
    cat[:, h:2 * h] = c_term
    cat[:, 2 * h:] = upper
    h1 = _e(b, h, device=dev)
    ops.gemm_nt(cat, w.wcomb, h1, b, h, 3 * h, 3 * h, 3 * h, h, bias=w.cbias, relu=True)
    h2 = _e(b, h, device=dev)
    ops.gemm_nt(h1, w.w2, h2, b, h, h, h, h, h, bias=w.b2, relu=True)
    lse = _e(b, dtype=F32, device=dev)
    lpt = _e(b, dtype=F32, device=dev)
    logp = _e(b, q, dtype=F32, device=dev)
    tgt = torch.zeros(b, dtype=torch.uint8, device=dev)
    ops.gemm_nll(1, h2, w.w3, w.b3, tgt, b, h, h, h, lse=lse, logp_target=lpt, logp=logp)
    return logp


#: set True to decode greedily (argmax) instead of sampling - lets eager and CUDA-graph runs be compared exactly
_GREEDY = False


def _frame_phase(p, tiers, sw, lut, win, conds_cur, c_term, outs, states, frame_out, logp_frame, generator):
    """One sample step at phase ``p = xi % FS`` of a top-tier frame, written against STATIC buffers so that it can be
    captured in a CUDA graph: ``win`` (B,FS) holds the last FS generated samples, ``outs`` the tiers' current
    upsampled outputs, ``frame_out[:, p]`` receives the new sample."""
    fs_top = win.shape[1]
    for n in reversed(range(len(tiers))):                                    # model.py:312-337
        w = tiers[n]
        if p % w.fs != 0:
            continue
        upper = None
        if n != len(tiers) - 1:
            frame_index = (p % tiers[n + 1].fs) // w.fs                      # == (xi // fs_n) % r_{n+1}
            upper = outs[n + 1][:, frame_index].contiguous()
        up = tier_step(w, lut, win[:, fs_top - w.fs:].contiguous(), conds_cur, upper, states[n])
        if outs[n] is None:
            outs[n] = up
        else:
            outs[n].copy_(up)
    upper = outs[0][:, p % sw.r0].contiguous()                               # model.py:343
    logp = sample_step(sw, win[:, fs_top - sw.r0:].contiguous(), c_term, upper)
    if logp_frame is not None:
        logp_frame[:, p] = logp
    if _GREEDY:                                                              # debugging aid: deterministic decoding
        new = logp.argmax(dim=1, keepdim=True).to(torch.uint8)
    else:
        new = torch.multinomial(logp.exp(), 1, generator=generator).to(torch.uint8)   # model.py:346-348
    frame_out[:, p] = new[:, 0]
    win.copy_(torch.cat([win[:, 1:], new], dim=1))


@torch.no_grad()
def generate(model, utt_conds, info, return_logp=False, generator=None, use_graphs=True):
    """model.py:289-351, batched over the first dimension of ``utt_conds`` (the reference handles one
    utterance per call).  Returns int64 (B, (t+1)*FS) with FS leading ``quantize_zero()`` samples; with
    ``return_logp`` also the (B, t*FS, Q) log-probabilities each sample was drawn from.

    The FS sample steps of one top-tier frame always run the same kernels on the same buffers (which tier
    fires and which upsampled vector is read depends only on ``xi % FS``), so after an eager first frame the
    FS step programs are captured once as CUDA graphs and replayed for every later frame: the launch-bound
    Python loop (~300 us per sample step) becomes FS graph replays per frame."""
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
    if any(layer.rnn_cell != 'gru' for layer in model.frames_layers):
        raise NotImplementedError('generation is implemented for GRU tiers (the reference cell)')
    tiers = [TierWeights(layer, c) for layer in model.frames_layers]
    sw = SampleWeights(model.sample_layer, c)
    # learnable h0 (model.py:111); clone(): for b == 1 expand().contiguous() would alias the parameter itself
    states = [w.h0[:, None, :].expand(-1, b, -1).clone() for w in tiers]
    outs = [None] * len(tiers)
    logps = [] if return_logp else None
    cp = sw.cp
    win = y[:, :fs_top].clone()                                              # the FS samples before the frame
    conds_cur = torch.empty(b, 1, c, dtype=F32, device=dev)
    conds_b = _e(b, cp, device=dev)
    c_term = _e(b, sw.h, device=dev)
    frame_out = torch.empty(b, fs_top, dtype=torch.uint8, device=dev)
    logp_frame = torch.empty(b, fs_top, sw.q, dtype=F32, device=dev) if return_logp else None
    graphs = None
    # a custom generator cannot be captured; above ~128 utterances the step is GPU-bound and the graph's extra
    # buffer copies cost more than the launch overhead they save (measured: B=64 158 vs 280 us/step, B=256 358 vs 314)
    graphed = use_graphs and generator is None and t > 2 and b <= 128
    for f in range(t):                                                       # top-tier frames; xi = (f+1)*FS + p
        conds_cur.copy_(conds[:, f: f + 1])                                  # model.py:308-309: conds index xi//FS - 1
        ops.pad_cast_bf16(conds_cur.view(b, c), b, c, c, conds_b, cp, cp)
        ops.gemm_nt(conds_b, sw.wcs, c_term, b, sw.h, cp, cp, cp, sw.h, bias=sw.csb)
        if graphed and f == 1:                                               # frame 0 ran eagerly (lazy init done)
            graphs = []
            pool = None
            torch.cuda.synchronize()
            for p in range(fs_top):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    _frame_phase(p, tiers, sw, lut, win, conds_cur, c_term, outs, states, frame_out, logp_frame, None)
                pool = g.pool()
                graphs.append(g)
        if graphs is not None:
            for g in graphs:
                g.replay()
        else:
            for p in range(fs_top):
                _frame_phase(p, tiers, sw, lut, win, conds_cur, c_term, outs, states, frame_out, logp_frame, generator)
        y[:, (f + 1) * fs_top: (f + 2) * fs_top] = frame_out
        if return_logp:
            logps.append(logp_frame.clone())
    out = y.to(torch.int64)
    if return_logp:
        return out, torch.cat(logps, dim=1)
    return out
