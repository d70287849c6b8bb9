"""The fp32-tolerance arithmetic mode (``SampleRNNModel(precision='fp32')``): the same decomposition of the reference's
three modules as ``functional.py`` (folded embedding table, conditioning hoisted to frame rate, hand-written backward
passes; model.py:28-203), but every activation, gate and gradient is an fp32 tensor and every contraction enters the
tcgen05 GEMM on split-bf16 operands (``ops.gemm_nt32`` / ``ops.gemm_tn32``: a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo in
one GEMM with a 3x longer K, fp32 accumulation), so the results agree with the reference's fp32 arithmetic to fp32-level
tolerances (loss rel <= 1e-5, per-tensor gradient rel-L2 <= 3e-3).  The tensor core's fp32 accumulator
truncates (~3e-8 relative to the accumulator per K=16 update, measured in tests/test_gpu_fp32_mode.py), so the segments are
concatenated SMALLEST PRODUCT FIRST: the leading hi.hi product comes last and only its K/16 updates truncate at full
magnitude.  FORWARD contractions use three pieces per operand and six products (``terms=6``: ~1e-6 at K = 1024, 8e-8 at
short K), because what a forward error costs is not linear: a pre-activation that lands on the other side of zero flips a
ReLU gate, and every flipped gate is a full-size error in the gradients behind it.  Backward contractions (linear in the
error) use two pieces and three products (4.5e-6).  The recurrence runs as one split-operand GEMM (six products forward,
three backward) and one fp32 cell kernel per timestep (``srnn_gru_forward_f32``).  GRU tiers only.  About 5-6x slower than
the bf16 path; it exists for validation, not for throughput.
"""
import torch

from . import ops
from .ops import F32, round_up


def _zeros(*shape, device=None):
    return torch.zeros(*shape, dtype=F32, device=device)


def _empty(*shape, device=None):
    return torch.empty(*shape, dtype=F32, device=device)


# ----------------------------------------------------------------------------------------------
# CondsMixer (model.py:60-65)
# ----------------------------------------------------------------------------------------------
class CondsMixFn32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, utt, spk_ids, table, weight, bias):
        b, l, u = utt.shape
        s = table.shape[1]
        c = weight.shape[0]
        kp = round_up(s + u, 8)
        utt = utt.contiguous()
        mixin = ops.mixer_input_f32(utt, table.contiguous(), spk_ids, kp)          # (B*L, kp)
        w = weight.contiguous()                                                    # (C, S+U): already K-major
        conds = ops.gemm_nt32(mixin[:, :s + u], w, bias=bias.contiguous(), terms=6)
        ctx.save_for_backward(mixin, w, spk_ids)
        ctx.dims = (b, l, u, s, c, kp)
        ctx.table_rows = table.shape[0]
        return conds.view(b, l, c)

    @staticmethod
    def backward(ctx, dconds):
        mixin, w, spk_ids = ctx.saved_tensors
        b, l, u, s, c, kp = ctx.dims
        dev = dconds.device
        dc = dconds.contiguous().view(b * l, c)
        dw = ops.gemm_tn32(dc, mixin[:, :s + u], _zeros(c, s + u, device=dev))
        dbias = ops.colsum_f32(dc)
        dmix = ops.gemm_nt32(dc, w.t().contiguous())                               # (B*L, S+U)
        dtable = _zeros(ctx.table_rows, s, device=dev)
        ops.mixer_input_bwd_f32(dmix, spk_ids, b, l, s, s + u, dtable)
        dutt = dmix[:, s:].contiguous().view(b, l, u) if ctx.needs_input_grad[0] else None
        return dutt, None, dtable, dw, dbias


# ----------------------------------------------------------------------------------------------
# FrameLevelLayer (model.py:140-156)
# ----------------------------------------------------------------------------------------------
class FrameTierFn32(torch.autograd.Function):
    """Same calling convention as ``functional.FrameTierFn`` with fp32 ``upper`` / outputs; GRU only."""

    @staticmethod
    def forward(ctx, xq_u8, x_off, lut, frames, conds, upper, h_init, fs, ratio, into_cat, c_init,
                xg, xv, xb, cg, cv, cb, ug, uv, ub, *rnn):
        if c_init is not None:
            raise NotImplementedError("precision='fp32' supports GRU tiers only")
        dev = conds.device
        b, l, c = conds.shape
        layers, _, h = h_init.shape
        if frames is not None:
            frames = frames.contiguous()
            t = frames.shape[1]
        else:
            t = (xq_u8.shape[1] - x_off) // fs if upper is None else upper.shape[1]
        r = ratio
        kp = round_up(fs + c, 8)
        conds = conds.contiguous()
        ain = ops.tier_input_f32(xq_u8, x_off, lut, frames, conds, b, t, fs, kp)

        wcat = _zeros(h, kp, device=dev)                                           # [Wx | Wc | 0], weight-normed
        inv_x = _empty(h, device=dev)
        inv_c = _empty(h, device=dev)
        ops.weight_prep_f32(xv, xg, (h, fs, 1), wcat, (kp, 1, 0), inv_norm=inv_x)
        ops.weight_prep_f32(cv, cg, (h, c, 1), wcat[:, fs:], (kp, 1, 0), inv_norm=inv_c)
        u = ops.gemm_nt32(ain, wcat, bias=(xb + cb), terms=6)
        if upper is not None:
            ops.bias_act_f32(u, aux2=upper.reshape(b * t, h))

        saved_layers = []
        x_l = u
        hn = _empty(layers, b, h, device=dev)
        for i in range(layers):
            w_ih, w_hh, b_ih, b_hh = (p.contiguous() for p in rnn[4 * i: 4 * i + 4])
            gi = ops.gemm_nt32(x_l, w_ih, bias=b_ih, terms=6)
            h0_i = h_init[i].contiguous()
            h_state = h0_i.clone()
            hall, gates = ops.gru_forward_f32(gi, w_hh, b_hh, h_state, b, t, h)
            hn[i] = h_state
            saved_layers.append((w_ih, w_hh, h0_i, hall, gates, x_l))
            x_l = hall

        wu = _empty(r * h, h, device=dev)                                          # wu[j*H+o, i] = Wu[i, o, j]
        wu_t = _empty(h, r * h, device=dev)
        inv_u = _empty(h, device=dev)
        ops.weight_prep_f32(uv, ug, (h, h, r), wu, (1, h, h * h), wu_t, (r * h, 1, h), inv_norm=inv_u)
        up = ops.gemm_nt32(x_l, wu, bias=ub.t().contiguous().view(-1), terms=6).view(b, t * r, h)

        ctx.dims = (b, t, l, c, h, fs, r, kp, layers, upper is not None)
        ctx.saved_layers = saved_layers
        ctx.save_for_backward(ain, wcat, wu_t, inv_x, inv_c, inv_u, xg, xv, cg, cv, ug, uv)
        cn = _empty(0, device=dev)
        ctx.mark_non_differentiable(hn, cn)
        return up, hn, cn

    @staticmethod
    def backward(ctx, dup, _dhn, _dcn):
        ain, wcat, wu_t, inv_x, inv_c, inv_u, xg, xv, cg, cv, ug, uv = ctx.saved_tensors
        b, t, l, c, h, fs, r, kp, layers, has_upper = ctx.dims
        dev = dup.device
        dup = dup.contiguous().view(b * t, r * h)
        last_hall = ctx.saved_layers[-1][3]
        d_ub = ops.colsum_f32(dup).view(r, h).t().contiguous()
        dwu = ops.gemm_tn32(dup, last_hall, _zeros(r * h, h, device=dev))
        d_uv, d_ug = ops.weight_prep_bwd(dwu, (1, h, h * h), uv, ug, inv_u, (h, h, r))
        dh_out = ops.gemm_nt32(dup, wu_t)                                          # (B*T, H)

        rnn_grads = [None] * (4 * layers)
        dh0 = _empty(layers, b, h, device=dev)
        d_bias = None
        for i in reversed(range(layers)):
            w_ih, w_hh, h0_i, hall, gates, x_l = ctx.saved_layers[i]
            dgi, dgh, dh0_i = ops.gru_backward_f32(w_hh, gates, hall, h0_i, dh_out, b, t, h)
            dh0[i] = dh0_i
            hprev = _empty(b, t, h, device=dev)                                    # h_{t-1} per row (device copies only)
            hprev[:, 0] = h0_i
            if t > 1:
                hprev[:, 1:] = hall.view(b, t, h)[:, :-1]
            dwhh = ops.gemm_tn32(dgh, hprev.view(b * t, h), _zeros(3 * h, h, device=dev))
            dwih = ops.gemm_tn32(dgi, x_l, _zeros(3 * h, h, device=dev))
            rnn_grads[4 * i: 4 * i + 4] = [dwih, dwhh, ops.colsum_f32(dgi), ops.colsum_f32(dgh)]
            if i == 0:
                d_bias = _zeros(h, device=dev)
            dh_out = ops.gemm_nt32(dgi, w_ih.t().contiguous(), colsum=d_bias if i == 0 else None)
        du = dh_out
        dwcat = ops.gemm_tn32(du, ain, _zeros(h, kp, device=dev))
        d_xv, d_xg = ops.weight_prep_bwd(dwcat, (kp, 1, 0), xv, xg, inv_x, (h, fs, 1))
        d_cv, d_cg = ops.weight_prep_bwd(dwcat[:, fs:], (kp, 1, 0), cv, cg, inv_c, (h, c, 1))
        dconds = None
        if ctx.needs_input_grad[4]:
            dc_rows = ops.gemm_nt32(du, wcat[:, fs:fs + c].t().contiguous())       # (B*T, C)
            dconds = _zeros(b, l, c, device=dev)
            ops.tier_input_bwd_f32(dc_rows, b, t, 0, l, c, dconds)
        d_upper = du.view(b, t, h) if has_upper else None
        return (None, None, None, None, dconds, d_upper, dh0, None, None, None, None,
                d_xg.view_as(xg), d_xv, d_bias, d_cg.view_as(cg), d_cv, d_bias.clone(),
                d_ug.view_as(ug), d_uv, d_ub, *rnn_grads)


# ----------------------------------------------------------------------------------------------
# SampleLevelLayer (model.py:188-203) + the NLL of runner.py:52
# ----------------------------------------------------------------------------------------------
class SampleLevelFn32(torch.autograd.Function):
    """Same decomposition as ``functional.SampleLevelFn`` (one-hot x folded table, conditioning block at frame rate)."""

    @staticmethod
    def forward(ctx, xs_u8, conds, upper, target_u8, fused, emb, eg, ev, csw, csb, cw, cbias, w2g, w2v, b2, w3g, w3v, b3):
        dev = conds.device
        b, w = xs_u8.shape
        _, l, c = conds.shape
        h, q, r0 = ev.shape
        rf = w - r0 + 1
        m = b * rf
        fsz = rf // l

        onehot = ops.onehot_rows(xs_u8, q)                                         # (B, W, Q) bf16, exact
        e32 = emb.contiguous()
        we = _empty(h, r0 * q, device=dev)                                         # We[o, k*Q+q']
        we_t = _empty(q, r0 * h, device=dev)                                       # We^T[q', k*H+o]
        inv_e = _empty(h, device=dev)
        ops.weight_prep_f32(ev, eg, (h, q, r0), we, (r0 * q, 1, q), we_t, (1, r0 * h, h), inv_norm=inv_e)
        tt = _empty(r0 * q, h, device=dev)                                         # tt[k*Q+q, o] = sum_q' E[q,q'] We[o,q',k]
        for k in range(r0):
            ops.gemm_nt32(e32, we[:, k * q:(k + 1) * q], out=tt[k * q:(k + 1) * q], terms=6)
        cwc = cw.contiguous()
        w_e, w_c, w_u = cwc[:, :h], cwc[:, h:2 * h], cwc[:, 2 * h:]
        tprime_t = ops.gemm_nt32(tt, w_e, terms=6)                                 # T'^T[k*Q+q, o'] = sum_o tt[k*Q+q,o] W_e[o',o]

        conds2 = conds.contiguous().view(b * l, c)
        c_frame = ops.gemm_nt32(conds2, csw.contiguous().view(h, c), bias=csb.contiguous(), terms=6)
        cterm = ops.gemm_nt32(c_frame, w_c, bias=cbias.contiguous(), terms=6)      # (B*L, H)

        # the one-hot x table product is a gather-sum of r0 table rows per sample: exact in fp32
        p_e = ops.embed_gather_f32(tprime_t, xs_u8.contiguous(), rf, r0, q)        # (m, H)
        upper_c = upper.reshape(m, h).contiguous()
        h1 = ops.gemm_nt32(upper_c, w_u, terms=6)
        mk1 = torch.empty(m, (h + 31) // 32, dtype=torch.int32, device=dev)
        mk2 = torch.empty(m, (h + 31) // 32, dtype=torch.int32, device=dev)
        ops.bias_act_f32(h1, aux=cterm, aux_row_div=fsz, aux2=p_e, relu=True, mask=mk1)
        del p_e

        w2 = _empty(h, h, device=dev)
        w2_t = _empty(h, h, device=dev)
        inv_2 = _empty(h, device=dev)
        ops.weight_prep_f32(w2v, w2g, (h, h, 1), w2, (h, 1, 0), w2_t, (1, h, 0), inv_norm=inv_2)
        h2 = ops.gemm_nt32(h1, w2, bias=b2.contiguous(), terms=6)
        ops.bias_act_f32(h2, relu=True, mask=mk2)
        w3 = _empty(q, h, device=dev)
        w3_t = _empty(h, q, device=dev)
        inv_3 = _empty(q, device=dev)
        ops.weight_prep_f32(w3v, w3g, (q, h, 1), w3, (h, 1, 0), w3_t, (1, q, 0), inv_norm=inv_3)
        logp = ops.gemm_nt32(h2, w3, bias=b3.contiguous(), terms=6)                # logits, then log-probabilities in place
        if target_u8 is None:
            target_u8 = torch.zeros(m, dtype=torch.uint8, device=dev)
        target_u8 = target_u8.contiguous()
        _, logp_t = ops.logsoftmax_nll_f32(logp, target_u8)
        out = logp_t.view(b, rf) if fused else logp.view(b, rf, q)
        ctx.dims = (b, w, l, c, h, q, r0, rf, m, fsz, fused)
        ctx.save_for_backward(onehot, e32, we_t, inv_e, conds2, csw, c_frame, upper_c, tt, cwc, h1, w2_t, inv_2, h2, w3_t,
                              inv_3, target_u8, logp, eg, ev, w2g, w2v, w3g, w3v, mk1, mk2)
        return out

    @staticmethod
    def backward(ctx, gout):
        (onehot, e32, we_t, inv_e, conds2, csw, c_frame, upper_c, tt, cwc, h1, w2_t, inv_2, h2, w3_t, inv_3, target_u8,
         logp, eg, ev, w2g, w2v, w3g, w3v, mk1, mk2) = ctx.saved_tensors
        b, w, l, c, h, q, r0, rf, m, fsz, fused = ctx.dims
        dev = gout.device
        gout = gout.contiguous().float()
        if fused:
            dlog = ops.logsoftmax_nll_bwd_f32(logp, target_u8, row_grad=gout.view(-1))
        else:
            dlog = ops.logsoftmax_nll_bwd_f32(logp, target_u8, g=gout.view(m, q))
        # adapt
        d_b3 = ops.colsum_f32(dlog)
        dw3 = ops.gemm_tn32(dlog, h2, _zeros(q, h, device=dev))
        d_w3v, d_w3g = ops.weight_prep_bwd(dw3, (h, 1, 0), w3v, w3g, inv_3, (q, h, 1))
        d_b2 = _zeros(h, device=dev)
        dh2 = ops.gemm_nt32(dlog, w3_t, gate_mask=mk2, colsum=d_b2)
        # comb_layer_expand
        dw2 = ops.gemm_tn32(dh2, h1, _zeros(h, h, device=dev))
        d_w2v, d_w2g = ops.weight_prep_bwd(dw2, (h, 1, 0), w2v, w2g, inv_2, (h, h, 1))
        d_cbias = _zeros(h, device=dev)
        dh1 = ops.gemm_nt32(dh2, w2_t, gate_mask=mk1, colsum=d_cbias)
        # comb_layer: [e | upper] blocks at sample rate, conditioning block at frame rate
        wcomb_t = cwc.t().contiguous()                                             # (3H, H): e | c | upper blocks of W^T
        d_cw = _zeros(h, 3 * h, device=dev)
        ops.gemm_tn32(dh1, upper_c, d_cw[:, 2 * h:])                               # d W_u
        seg = ops.segment_sum_f32(dh1, fsz)                                        # (B*L, H)
        ops.gemm_tn32(seg, c_frame, d_cw[:, h:2 * h])                              # d W_c
        dupper = ops.gemm_nt32(dh1, wcomb_t[2 * h:])
        dc_frame = ops.gemm_nt32(seg, wcomb_t[h:2 * h])
        # conds_expand (frame rate)
        d_csb = ops.colsum_f32(dc_frame)
        dwcs = ops.gemm_tn32(dc_frame, conds2, _zeros(h, c, device=dev))
        dconds = ops.gemm_nt32(dc_frame, csw.contiguous().view(h, c).t().contiguous())
        # folded table: gt[k*Q+q, o'] = sum_j window_j[k*Q+q] dh1[j, o'] - the one-hot side is exact, dh1 enters as hi + lo
        d2, hp = ops.split3(dh1, 2)                                                # bf16 (m, 2*hp)
        gt = _zeros(r0 * q, h, device=dev)
        for i in range(2):
            ops.gemm_tn(onehot, d2[:, i * hp:], gt, r0 * q, h, rf, q, 2 * hp, h, batch=b, a_bs=w * q, b_bs=rf * 2 * hp)
        ops.gemm_tn32(gt, tt, d_cw[:, :h])                                         # d W_e[o',o] = sum gt[.,o'] tt[.,o]
        g = _empty(q, r0 * h, device=dev)                                          # G[q, k*H+o] = sum_o' gt[kQ+q,o'] W_e[o',o]
        for k in range(r0):
            ops.gemm_nt32(gt[k * q:(k + 1) * q], wcomb_t[:h], out=g[:, k * h:(k + 1) * h])
        d_emb = ops.gemm_nt32(g, we_t)                                             # d E[q,q'] = sum G[q,kH+o] We[o,q',k]
        dwe = _zeros(h, r0 * q, device=dev)
        for k in range(r0):
            ops.gemm_tn32(g[:, k * h:(k + 1) * h], e32, dwe[:, k * q:(k + 1) * q])
        d_ev, d_eg = ops.weight_prep_bwd(dwe, (r0 * q, 1, q), ev, eg, inv_e, (h, q, r0))
        return (None, dconds.view(b, l, c), dupper.view(b, rf, h), None, None,
                d_emb, d_eg.view_as(eg), d_ev, dwcs.view(h, c, 1), d_csb, d_cw, d_cbias,
                d_w2g.view_as(w2g), d_w2v, d_b2, d_w3g.view_as(w3g), d_w3v, d_b3)
