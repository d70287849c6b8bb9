"""Autoregressive generation (``SampleRNNModel.test``, model.py:289-351; BASELINE config 5).

Batched, sample-by-sample generation on the same C-ABI kernels as training.  Generation is a chain of small dependent
launches, so what matters is the LENGTH of that chain per sample; with fixed weights most of it can be composed ahead
of time (once per call):

  * tier input projection + ``W_ih``:  gi = W_ih (W_in a + b_in + upper) + b_ih = (W_ih W_in) a + W_ih upper + const.
    ``W_ih W_in`` is a (3H x ~56) matrix; ``W_ih upper`` is produced by the tier ABOVE at its own (slower) rate, because
    ``upper`` is itself linear in that tier's hidden state: W_ih (W_up,j h + b_j) = (W_ih W_up,j) h + const;
  * learned upsampling of the lowest tier + the upper-tier block of ``comb_layer``:  pre_j = (W_cu W_up,j) h + const;
  * the embedding + conv1d + the embedding block of ``comb_layer`` are r0 tables of 256 rows (as in training);
  * the bf16 copy of h_t that the recurrent kernel writes is the next step's MMA operand (no cast kernel in between);
  * the draw of sample i and the table gather of sample i+1 are one launch (``srnn_sample_embed``).

Per frame of a tier that leaves: operand assembly, one K~56 GEMM, the recurrent step (one launch for all utterances),
one K=H GEMM; per sample: comb_layer_expand, adapt, draw(+gather).  The FS step programs of a top-tier frame are captured
once as ONE CUDA graph and replayed; the draws are device-side Philox numbers, so the graph holds the whole sample step.
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
        self.wcat_t = _z(self.kp, h, device=dev)                       # [W_x | W_c]^T for the composition with W_ih
        ops.weight_prep(layer.x_expand.weight_v.detach(), layer.x_expand.weight_g.detach(), (h, fs, 1), self.wcat,
                        (self.kp, 1, 0), self.wcat_t, (1, h, 0))
        ops.weight_prep(layer.conds_expand.weight_v.detach(), layer.conds_expand.weight_g.detach(), (h, c, 1),
                        self.wcat[:, fs:], (self.kp, 1, 0), self.wcat_t[fs:], (1, h, 0))
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
        self.wu = _e(r * h, h, device=dev)                              # rows (j, o): W_up,j[o, i]
        self.wu_t = _e(r * h, h, device=dev)                            # rows (j, i): W_up,j[o, i] transposed
        ops.weight_prep(layer.upsample.weight_v.detach(), layer.upsample.weight_g.detach(), (h, h, r), self.wu,
                        (1, h, h * h), self.wu_t, (h, 1, h * h))
        self.bias_up = layer.upsample_bias.detach().t().contiguous().view(-1)
        self.h0 = layer.rnn_h0.detach()
        self.c0 = layer.rnn_c0.detach() if self.lstm else None
        # composed first-layer input weights: gi = a . wa^T + ba (+ the upper tier's contribution)
        wih0, _, b_ih0, _ = self.rnn[0]
        self.wa = _e(ng * h, self.kp, device=dev)
        ops.gemm_nt(wih0, self.wcat_t, self.wa, ng * h, self.kp, h, h, h, self.kp)
        self.ba = _bias_through(wih0, self.bias_u.view(1, h), ng * h, h)[0] + b_ih0
        self.out_w = self.out_b = None                                  # composed output GEMM, set by compose_outputs()
        self.out_n = 0

    def compose_output(self, w_next, b_extra, n_per):
        """out_w[(j, g), :] = (w_next W_up,j)[g, :] for the r sub-frames j, ``w_next`` (n_per x H) bf16 being what the
        consumer applies to this tier's upsampled output (the tier below: its W_ih; the sample level: the upper block of
        comb_layer); out_b = w_next b_up,j + b_extra."""
        h, r, dev = self.h, self.r, self.wu.device
        self.out_n = r * n_per
        self.out_w = _e(r * n_per, h, device=dev)
        lda = w_next.stride(0)
        for j in range(r):
            ops.gemm_nt(w_next, self.wu_t[j * h:], self.out_w[j * n_per:], n_per, h, h, lda, h, h)
        self.out_b = (_bias_through(w_next, self.bias_up.view(r, h), n_per, h, lda) + b_extra).reshape(-1).contiguous()


def _bias_through(w, rows, n, k, ldw=None):
    """fp32 (len(rows), n) = rows . w^T for a few bias vectors ``rows`` (fp32 (m, k)) through the tensor-core GEMM (m is
    padded to 8 rows); used once per call for the composed biases."""
    m = rows.shape[0]
    a = torch.zeros(round_up(m, 8), k, dtype=F32, device=w.device)
    a[:m] = rows
    out = torch.empty(round_up(m, 8), n, dtype=F32, device=w.device)
    ops.gemm_nt(ops.to_bf16(a), w, out, round_up(m, 8), n, k, k, ldw if ldw is not None else k, n)
    return out[:m]


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


def tier_step(w, lut, win, conds_row, aux, aux_ld, h_state, hx, step, out, out_aux, c_state=None):
    """One frame of a tier for every utterance, composed form.  ``win`` (B, FS) uint8 holds the last FS generated samples
    (the tier reads its last fs), ``conds_row`` (B,1,C) fp32, ``aux`` a (B, ng*H) bf16 view with row stride ``aux_ld``
    holding the upper tier's contribution to the gate pre-activations (None for the top tier: the constant is then the
    GEMM bias), ``h_state`` (and ``c_state``) (layers,B,H) fp32 updated in place, ``hx`` the per-layer chains of bf16
    hidden states (slot ``step`` is read, ``step + 1`` written).  Writes ``out`` (B, w.out_n) bf16: what the consumer
    below needs for the r sub-frames (see TierWeights.compose_output)."""
    b, fs_top = win.shape
    dev = win.device
    h = w.h
    ain = ops.tier_input(win, fs_top - w.fs, lut, None, conds_row, b, 1, w.fs, w.kp)
    gi = _e(b, w.ng * h, device=dev)
    ops.gemm_nt(ain, w.wa, gi, b, w.ng * h, w.kp, w.kp, w.kp, w.ng * h, bias=w.ba if aux is None else None, aux=aux,
                ldaux=aux_ld, aux_mode=1)
    x = None
    for i, (wih, whh, b_ih, b_hh) in enumerate(w.rnn):
        if i > 0:
            gi = _e(b, w.ng * h, device=dev)
            ops.gemm_nt(x, wih, gi, b, w.ng * h, h, h, h, w.ng * h, bias=b_ih)
        h_ext = hx[i][step: step + 2]                               # slot `step` = h_{t-1} in bf16 (the previous step's output)
        hall = _e(b, h, device=dev)
        if w.lstm:
            gates = _e(b, 5 * h, device=dev)
            ops.lstm_forward(gi, whh, b_hh, h_ext, hall, h_state[i], c_state[i], gates, b, 1, h)
        else:
            gates = _e(b, 4 * h, device=dev)
            ops.gru_forward(gi, whh, b_hh, h_ext, hall, h_state[i], gates, b, 1, h)
        x = hall
    ops.gemm_nt(x, w.out_w, out, b, w.out_n, h, h, h, w.out_n, bias=w.out_b, aux=out_aux, ldaux=w.out_n,
                aux_mode=1 if out_aux is not None else 0)
    return out


def frame_terms(sw, conds_b, c_term, cc, cc_rep):
    """Per top-tier frame: ``c_term`` = conds_expand(conds) (model.py:194) and ``cc`` = the conditioning block of
    comb_layer applied to it, plus comb_layer's bias (model.py:195-200) - both constant over the frame; ``cc_rep``
    repeats it for the r0 sub-frame column blocks of the lowest tier's composed output GEMM."""
    b, h, cp = conds_b.shape[0], sw.h, sw.cp
    ops.gemm_nt(conds_b, sw.wcs, c_term, b, h, cp, cp, cp, h, bias=sw.csb)
    ops.gemm_nt(c_term, sw.wcomb[:, h:], cc, b, h, h, h, 3 * h, h, bias=sw.cbias)
    cc_rep.view(b, sw.r0, h).copy_(cc[:, None, :])


def sample_step(sw, h1, logits):
    """Logits (B,256) fp32 of the next sample from the embedding-side activation ``h1`` (the sum of r0 table rows + the
    frame term, ReLU): comb_layer_expand and adapt (model.py:201-202); the log-softmax (model.py:203) is taken by the
    sampling kernel."""
    b, h, q = h1.shape[0], sw.h, sw.q
    h2 = _e(b, h, device=h1.device)
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
        self.outs = [_e(b, w.out_n, device=dev) for w in tiers]            # composed outputs (see compose_output)
        self.cc_rep = _e(b, sw.r0 * sw.h, device=dev)
        self.h1 = _e(b, sw.h, device=dev)                                  # embedding-side activation of the next step
        # chains of bf16 hidden states per tier and layer: a tier runs FS / fs_n steps per top-tier frame
        self.hx = [[_e(fs_top // w.fs + 1, b, w.h, device=dev) for _ in range(w.layers)] for w in tiers]
        # staging for GRAPH_FRAMES top-tier frames: one CUDA graph covers that many frames, the host refills / drains
        # the stages once per replay (a replay per frame left the host as busy as the GPU: ~0.5 ms of enqueue work per
        # 0.54 ms frame, so every host hiccup showed up in the throughput)
        self.conds_stage = torch.empty(b, GRAPH_FRAMES, c, dtype=F32, device=dev)
        self.out_stage = torch.empty(b, GRAPH_FRAMES * fs_top, dtype=torch.uint8, device=dev)
        self.logp_stage = torch.empty(b, GRAPH_FRAMES * fs_top, sw.q, dtype=F32, device=dev) if return_logp else None
        self.logits = torch.empty(b, sw.q, dtype=F32, device=dev)
        self.rng_state = torch.zeros(3, dtype=torch.int64, device=dev)      # {seed, step, 0} of the device-side draws


def _frame_phase(p, tiers, sw, lut, st, states, cstates, slot):
    """One sample step at phase ``p = xi % FS`` of a top-tier frame, written against the static buffers of ``st`` so
    that it can be captured in a CUDA graph: ``st.win`` (B,FS) holds the last FS generated samples, ``st.outs`` the
    tiers' current composed outputs, ``st.out_stage[:, slot*FS + p]`` receives the new sample."""
    win = st.win
    b, fs_top = win.shape
    col = slot * fs_top + p
    out_ld = st.out_stage.shape[1]
    r0, h = sw.r0, sw.h
    fired = False
    for n in reversed(range(len(tiers))):                                    # model.py:312-337
        w = tiers[n]
        if p % w.fs != 0:
            continue
        fired = True
        aux, aux_ld = None, 0
        if n != len(tiers) - 1:
            up = tiers[n + 1]
            frame_index = (p % up.fs) // w.fs                                # == (xi // fs_n) % r_{n+1}
            n_per = w.ng * w.h
            aux, aux_ld = st.outs[n + 1][:, frame_index * n_per:], up.out_n
        tier_step(w, lut, win, st.conds_cur, aux, aux_ld, states[n], st.hx[n], p // w.fs, st.outs[n],
                  st.cc_rep if n == 0 else None, cstates[n])
    j = p % r0                                                               # model.py:343
    if fired or p == 0:
        # the frame term of this step has just been produced: gather the embedding side here
        ops.embed_sum(sw.table_t, win[:, fs_top - r0:], fs_top, b, r0, sw.q, h, st.outs[0][:, j * h:], r0 * h, True, st.h1, h)
    sample_step(sw, st.h1, st.logits)
    # model.py:203,346-348: log-softmax, draw from it, append to the window of the last FS samples; unless a tier step
    # comes first, the same launch also gathers the embedding side of the NEXT step
    logp_out = st.logp_stage[:, col] if st.logp_stage is not None else None
    rng = None if _GREEDY else st.rng_state
    nxt = p + 1
    if nxt < fs_top and nxt % tiers[0].fs != 0:
        ops.sample_embed(st.logits, b, sw.q, win, fs_top, st.out_stage[:, col], out_ld, sw.table_t, r0, h,
                         st.outs[0][:, (nxt % r0) * h:], r0 * h, st.h1, h, normalise=True, logp_out=logp_out, rng_state=rng)
    else:
        ops.sample_categorical(st.logits, b, sw.q, None, win, fs_top, st.out_stage[:, col], out_ld, normalise=True,
                               logp_out=logp_out, rng_state=rng)


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
    for n, w in enumerate(tiers):                                            # what each tier hands to its consumer
        if n == 0:                                                           # sample level: upper block of comb_layer
            w.compose_output(sw.wcomb[:, 2 * sw.h:], torch.zeros(sw.h, device=dev), sw.h)
        else:                                                                # the tier below: its W_ih (+ its constant)
            below = tiers[n - 1]
            w.compose_output(below.rnn[0][0], below.ba, below.ng * below.h)
    # learnable h0 (model.py:111); clone(): for b == 1 expand().contiguous() would alias the parameter itself
    states = [w.h0[:, None, :].expand(-1, b, -1).clone() for w in tiers]
    cstates = [w.c0[:, None, :].expand(-1, b, -1).clone() if w.lstm else None for w in tiers]   # LSTM extension
    st = _GenState(b, fs_top, tiers, sw, c, return_logp, dev)
    st.win = y[:, :fs_top].clone()                                           # the FS samples before the frame
    for n, w in enumerate(tiers):                                            # slot 0 of every chain: the initial state
        for i in range(w.layers):
            ops.pad_cast_bf16(states[n][i], b, w.h, w.h, st.hx[n][i][0], w.h, w.h)
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


#: top-tier frames per CUDA graph
GRAPH_FRAMES = 8


def _run_frame(slot, first, tiers, sw, lut, st, states, cstates):
    """All FS sample steps of one top-tier frame against stage slot ``slot`` (graph-capturable)."""
    b, c = st.conds_cur.shape[0], st.conds_cur.shape[2]
    fs_top = st.win.shape[1]
    st.conds_cur.copy_(st.conds_stage[:, slot: slot + 1])                    # model.py:308-309: conds index xi//FS - 1
    ops.pad_cast_bf16(st.conds_cur.view(b, c), b, c, c, st.conds_b, sw.cp, sw.cp)
    frame_terms(sw, st.conds_b, st.c_term, st.cc, st.cc_rep)
    if not first:                                                            # the chains wrap: slot 0 <- last slot
        for chain in st.hx:
            for hx in chain:
                hx[0].copy_(hx[-1])
    for p in range(fs_top):
        _frame_phase(p, tiers, sw, lut, st, states, cstates, slot)


def _generate_frames(model, st, states, cstates, tiers, sw, lut, conds, y, t, b, c, fs_top, graphed, generator, return_logp):
    logps = [] if return_logp else None

    def drain(f0, n):
        y[:, (f0 + 1) * fs_top: (f0 + 1 + n) * fs_top] = st.out_stage[:, :n * fs_top]
        if return_logp:
            logps.append(st.logp_stage[:, :n * fs_top].clone())

    def eager(f0, n):
        st.conds_stage[:, :n].copy_(conds[:, f0: f0 + n])
        for i in range(n):
            _run_frame(i, f0 + i == 0, tiers, sw, lut, st, states, cstates)
        drain(f0, n)

    eager(0, 1)                                                              # frame 0 runs eagerly (lazy initialisation)
    f = 1
    graph, per_graph = None, 0
    while f < t:
        n = min(GRAPH_FRAMES, t - f)
        if not graphed or n < GRAPH_FRAMES:
            eager(f, n)                                                      # tail shorter than one graph
            f += n
            continue
        st.conds_stage.copy_(conds[:, f: f + n])
        if graph is None:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()                                   # ONE graph = GRAPH_FRAMES x FS step programs
            before = ops.launch_count
            with torch.cuda.graph(graph):
                for i in range(n):
                    _run_frame(i, False, tiers, sw, lut, st, states, cstates)
            per_graph = ops.launch_count - before                            # kernels inside the captured graph
            ops.launch_count = before
        graph.replay()
        ops.launch_count += per_graph
        drain(f, n)
        f += n
    out = y.to(torch.int64)
    if return_logp:
        return out, torch.cat(logps, dim=1)
    return out
