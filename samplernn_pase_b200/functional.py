"""``torch.autograd.Function``s that compose the C-ABI kernels into the reference's three
modules (CondsMixer, FrameLevelLayer, SampleLevelLayer; model.py:28-203) with hand-written
backward passes.  Arithmetic: bf16 operands, fp32 accumulation (tcgen05), fp32 recurrent state,
fp32 parameters and gradients.
"""
import torch

from . import ops
from .ops import BF16, F32, round_up


def _zeros(*shape, dtype=F32, device=None):
    return torch.zeros(*shape, dtype=dtype, device=device)


def _empty(*shape, dtype=BF16, device=None):
    return torch.empty(*shape, dtype=dtype, device=device)


# ----------------------------------------------------------------------------------------------
# sum of the negative log-likelihoods (runner.py:52 with reduction='sum') on our own reduction kernel
# ----------------------------------------------------------------------------------------------
class NegSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp_target):
        flat = logp_target.contiguous().view(-1)
        ctx.shape = logp_target.shape
        if flat.numel() == 0:
            # a rank whose slots are all empty (reset == 2, the tail of an epoch): zero loss that still belongs to
            # the graph, so backward runs (with zero gradients) and the bucket all-reduces stay matched across ranks
            return flat.new_zeros(())
        one = torch.ones(1, dtype=torch.uint8, device=flat.device)
        out = ops.masked_nll_mean(flat, one, flat.numel())          # [-mean, count]
        return out[0] * out[1]

    @staticmethod
    def backward(ctx, g):
        return (-g).expand(ctx.shape).contiguous()


# ----------------------------------------------------------------------------------------------
# CondsMixer (model.py:60-65): conds = Linear([speaker_emb | utt])
# ----------------------------------------------------------------------------------------------
class CondsMixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, utt, spk_ids, table, weight, bias):
        b, l, u = utt.shape
        s = table.shape[1]
        c = weight.shape[0]
        dev = utt.device
        kp = round_up(s + u, 8)
        utt = utt.contiguous()
        mixin = ops.mixer_input(utt, table.contiguous(), spk_ids, kp)
        wb = _zeros(c, kp, dtype=BF16, device=dev)
        wbt = _zeros(kp, round_up(c, 8), dtype=BF16, device=dev)
        ops.weight_prep(weight.contiguous(), None, (c, s + u, 1), wb, (kp, 1, 0), wbt, (1, round_up(c, 8), 0))
        conds = _empty(b * l, c, dtype=F32, device=dev)
        ops.gemm_nt(mixin, wb, conds, b * l, c, kp, kp, kp, c, bias=bias.contiguous())
        ctx.save_for_backward(mixin, wbt, spk_ids)
        ctx.dims = (b, l, u, s, c, kp)
        ctx.table_rows = table.shape[0]
        return conds.view(b, l, c)

    @staticmethod
    def backward(ctx, dconds):
        mixin, wbt, spk_ids = ctx.saved_tensors
        b, l, u, s, c, kp = ctx.dims
        dev = dconds.device
        cp = round_up(c, 8)
        dcb = _empty(b * l, cp, device=dev)
        ops.pad_cast_bf16(dconds.contiguous(), b * l, c, c, dcb, cp, cp)
        dw = _zeros(c, kp, device=dev)
        ops.gemm_tn(dcb, mixin, dw, c, kp, b * l, cp, kp, kp)
        dbias = ops.colsum(dcb, b * l, c, cp)
        dmix = _empty(b * l, kp, device=dev)
        ops.gemm_nt(dcb, wbt, dmix, b * l, kp, cp, cp, cp, kp)
        dtable = _zeros(ctx.table_rows, s, device=dev)
        ops.mixer_input_bwd(dmix, spk_ids, b, l, s, kp, dtable)
        dutt = None
        if ctx.needs_input_grad[0]:
            dutt = _empty(b, l, u, dtype=F32, device=dev)
            ops.bf16_to_f32(dmix[:, s:], b * l, u, kp, dutt, u)
        return dutt, None, dtable, dw[:, :s + u].contiguous(), dbias


# ----------------------------------------------------------------------------------------------
# initial recurrent state (model.py:149-151 + 239-243)
# ----------------------------------------------------------------------------------------------
class StateSelectFn(torch.autograd.Function):
    """h_init[l, b] = carried[l, b] if use_carry[b] else rnn_h0[l]"""

    @staticmethod
    def forward(ctx, rnn_h0, carried, use_carry):
        layers, h = rnn_h0.shape
        b = use_carry.shape[0]
        outs = [ops.state_select(carried[i] if carried is not None else None, rnn_h0[i].contiguous(), use_carry, b, h)
                for i in range(layers)]
        ctx.save_for_backward(use_carry)
        return torch.stack(outs)

    @staticmethod
    def backward(ctx, dh):
        (use_carry,) = ctx.saved_tensors
        layers, b, h = dh.shape
        dh = dh.contiguous()
        return torch.stack([ops.state_select_bwd(dh[i], use_carry, b, h) for i in range(layers)]), None, None


# ----------------------------------------------------------------------------------------------
# FrameLevelLayer (model.py:140-156)
# ----------------------------------------------------------------------------------------------
class FrameTierFn(torch.autograd.Function):
    """inputs: (xq_u8, x_off, lut) or frames; conds (B,L,C) fp32; upper (B,T,H) bf16 or None;
    h_init (layers,B,H) fp32; then the tier parameters.  Returns (upsampled (B,T*r,H) bf16,
    h_n (layers,B,H) fp32, not differentiable - the reference detaches it, model.py:276)."""

    @staticmethod
    def forward(ctx, xq_u8, x_off, lut, frames, conds, upper, h_init, fs, ratio, into_cat, c_init,
                xg, xv, xb, cg, cv, cb, ug, uv, ub, *rnn):
        dev = conds.device
        b, l, c = conds.shape
        layers, _, h = h_init.shape
        lstm = c_init is not None                      # LSTM extension (BASELINE config 3), else GRU
        ng = 4 if lstm else 3
        if frames is not None:
            frames = frames.contiguous()
            t = frames.shape[1]
        else:
            t = (xq_u8.shape[1] - x_off) // fs if upper is None else upper.shape[1]
        r = ratio
        kp = round_up(fs + c, 8)
        cp = round_up(c, 8)
        conds = conds.contiguous()
        ain = ops.tier_input(xq_u8, x_off, lut, frames, conds, b, t, fs, kp)

        # input projections: [Wx | Wc] weight-normed, K-major; Wc^T kept for dconds
        wcat = _zeros(h, kp, dtype=BF16, device=dev)
        wct = _zeros(cp, h, dtype=BF16, device=dev)
        inv_x = _empty(h, dtype=F32, device=dev)
        inv_c = _empty(h, dtype=F32, device=dev)
        ops.weight_prep(xv, xg, (h, fs, 1), wcat, (kp, 1, 0), inv_norm=inv_x)
        ops.weight_prep(cv, cg, (h, c, 1), wcat[:, fs:], (kp, 1, 0), wct, (1, h, 0), inv_norm=inv_c)
        u = _empty(b * t, h, device=dev)
        if upper is not None:
            upper = upper.contiguous()
        ops.gemm_nt(ain, wcat, u, b * t, h, kp, kp, kp, h, bias=(xb + cb), aux=upper, ldaux=h, aux_mode=1)

        saved_layers = []
        x_l = u                                                # (B*T, H) batch-major input of layer 0
        hn = _empty(layers, b, h, dtype=F32, device=dev)
        cn = _empty(layers, b, h, dtype=F32, device=dev) if lstm else _empty(0, dtype=F32, device=dev)
        for i in range(layers):
            w_ih, w_hh, b_ih, b_hh = rnn[4 * i: 4 * i + 4]
            wih = _empty(ng * h, h, device=dev)
            wih_t = _empty(h, ng * h, device=dev)
            whh = _empty(ng * h, h, device=dev)
            whh_t = _empty(h, ng * h, device=dev)
            ops.weight_prep(w_ih.contiguous(), None, (ng * h, h, 1), wih, (h, 1, 0), wih_t, (1, ng * h, 0))
            ops.weight_prep(w_hh.contiguous(), None, (ng * h, h, 1), whh, (h, 1, 0), whh_t, (1, ng * h, 0))
            gi = _empty(b * t, ng * h, device=dev)
            ops.gemm_nt(x_l, wih, gi, b * t, ng * h, h, h, h, ng * h, bias=b_ih.contiguous())
            h_ext = _empty(t + 1, b, h, device=dev)            # time-major exchange buffer, slot 0 = h_init
            hall = _empty(b * t, h, device=dev)                # batch-major copy for the GEMMs
            h_state = h_init[i].contiguous().clone()
            ops.pad_cast_bf16(h_state, b, h, h, h_ext, h, h)
            c0_i = None
            if lstm:
                gates = _empty(b * t, 5 * h, device=dev)
                c0_i = c_init[i].contiguous()
                c_state = c0_i.clone()
                ops.lstm_forward(gi, whh, b_hh.contiguous(), h_ext, hall, h_state, c_state, gates, b, t, h)
                cn[i] = c_state
            else:
                gates = _empty(b * t, 4 * h, device=dev)
                ops.gru_forward(gi, whh, b_hh.contiguous(), h_ext, hall, h_state, gates, b, t, h)
            hn[i] = h_state
            saved_layers.append((wih_t, whh_t, h_ext, hall, gates, x_l, c0_i))
            x_l = hall

        # learned upsampling as one GEMM: out[(b,t), j*H+o] = sum_i h[b,t,i] Wu[i,o,j] + bias[o,j]
        wu = _empty(r * h, h, device=dev)
        wu_t = _empty(h, r * h, device=dev)
        inv_u = _empty(h, dtype=F32, device=dev)
        ops.weight_prep(uv, ug, (h, h, r), wu, (1, h, h * h), wu_t, (r * h, 1, h), inv_norm=inv_u)
        if into_cat and h % 32 == 0:
            # lowest tier: write the upsampled conditioning straight into column block [H,2H) of the
            # sample-level operand buffer [emb-conv | upper] (model.py:196-199 without the conditioning
            # block, which is hoisted to frame rate) - the returned tensor is a strided view of it
            cat = _empty(b * t * r, 2 * h, device=dev)
            up = cat[:, h:].view(b, t * r, h)
            ops.gemm_nt(x_l, wu, up, b * t, r * h, h, h, h, 2 * h, bias=ub.t().contiguous().view(-1), n_fold=h)
        else:
            up = _empty(b, t * r, h, device=dev)
            ops.gemm_nt(x_l, wu, up, b * t, r * h, h, h, h, r * h, bias=ub.t().contiguous().view(-1))

        ctx.dims = (b, t, l, c, h, fs, r, kp, cp, layers, upper is not None, lstm)
        ctx.saved_layers = saved_layers
        ctx.save_for_backward(ain, wct, wu_t, inv_x, inv_c, inv_u, xg, xv, cg, cv, ug, uv)
        ctx.mark_non_differentiable(hn, cn)
        return up, hn, cn

    @staticmethod
    def backward(ctx, dup, _dhn, _dcn):
        ain, wct, wu_t, inv_x, inv_c, inv_u, xg, xv, cg, cv, ug, uv = ctx.saved_tensors
        b, t, l, c, h, fs, r, kp, cp, layers, has_upper, lstm = ctx.dims
        ng = 4 if lstm else 3
        dev = dup.device
        dup = dup.contiguous()            # (B, T*r, H) == (B*T, r*H)
        last_hall = ctx.saved_layers[-1][3]
        # upsample
        d_ub = ops.colsum(dup, b * t, r * h, r * h).view(r, h).t().contiguous()
        dwu = _zeros(r * h, h, device=dev)
        ops.gemm_tn(dup, last_hall, dwu, r * h, h, b * t, r * h, h, h)
        d_uv, d_ug = ops.weight_prep_bwd(dwu, (1, h, h * h), uv, ug, inv_u, (h, h, r))
        dh_out = _empty(b * t, h, device=dev)
        ops.gemm_nt(dup, wu_t, dh_out, b * t, h, r * h, r * h, r * h, h)

        rnn_grads = [None] * (4 * layers)
        dh0 = _empty(layers, b, h, dtype=F32, device=dev)
        dc0 = _empty(layers, b, h, dtype=F32, device=dev) if lstm else None
        for i in reversed(range(layers)):
            wih_t, whh_t, h_ext, hall, gates, x_l, c0_i = ctx.saved_layers[i]
            dgi = _empty(b * t, ng * h, device=dev)            # batch-major
            dgh = _empty(t * b, ng * h, device=dev)            # time-major (exchange buffer)
            dh0_i = _empty(b, h, dtype=F32, device=dev)
            db_ih = _zeros(ng * h, device=dev)                 # bias gradients come out of the recurrent kernel itself
            db_hh = _zeros(ng * h, device=dev)
            if lstm:
                dc0_i = _empty(b, h, dtype=F32, device=dev)
                ops.lstm_backward(whh_t, h_ext, gates, c0_i, dh_out, dgi, dgh, dh0_i, dc0_i, b, t, h, db_ih, db_hh)
                dc0[i] = dc0_i
            else:
                ops.gru_backward(whh_t, h_ext, gates, dh_out, dgi, dgh, dh0_i, b, t, h, db_ih, db_hh)
            dh0[i] = dh0_i
            dwhh = _zeros(ng * h, h, device=dev)               # both operands time-major: rows (t, b)
            ops.gemm_tn(dgh, h_ext, dwhh, ng * h, h, t * b, ng * h, h, h)
            dwih = _zeros(ng * h, h, device=dev)
            ops.gemm_tn(dgi, x_l, dwih, ng * h, h, b * t, ng * h, h, h)
            rnn_grads[4 * i: 4 * i + 4] = [dwih, dwhh, db_ih, db_hh]
            dx = _empty(b * t, h, device=dev)
            d_bias = _zeros(h, dtype=F32, device=dev) if i == 0 else None   # layer 0: dx is du, its column sums the bias gradient
            ops.gemm_nt(dgi, wih_t, dx, b * t, h, ng * h, ng * h, ng * h, h, colsum=d_bias)
            dh_out = dx
        du = dh_out                       # (B*T, H): gradient of u, hence also of `upper`
        dwcat = _zeros(h, kp, device=dev)
        ops.gemm_tn(du, ain, dwcat, h, kp, b * t, h, kp, kp)
        d_xv, d_xg = ops.weight_prep_bwd(dwcat, (kp, 1, 0), xv, xg, inv_x, (h, fs, 1))
        d_cv, d_cg = ops.weight_prep_bwd(dwcat[:, fs:], (kp, 1, 0), cv, cg, inv_c, (h, c, 1))
        dconds = None
        if ctx.needs_input_grad[4]:
            dc_rows = _empty(b * t, cp, device=dev)
            ops.gemm_nt(du, wct, dc_rows, b * t, cp, h, h, h, cp)
            dconds = _zeros(b, l, c, device=dev)
            ops.tier_input_bwd(dc_rows, b, t, 0, l, c, cp, dconds)
        d_upper = du.view(b, t, h) if has_upper else None
        return (None, None, None, None, dconds, d_upper, dh0, None, None, None, dc0,
                d_xg.view_as(xg), d_xv, d_bias, d_cg.view_as(cg), d_cv, d_bias.clone(),
                d_ug.view_as(ug), d_uv, d_ub, *rnn_grads)


# ----------------------------------------------------------------------------------------------
# SampleLevelLayer (model.py:188-203) + the NLL of runner.py:52 when ``target`` is given
# ----------------------------------------------------------------------------------------------
class SampleLevelFn(torch.autograd.Function):
    """mode 'fused': returns log p(target) per row, (B,RF) fp32 (logits never reach HBM).
    mode 'full' : returns the (B,RF,Q) log-probabilities like the reference module.

    comb_layer (model.py:195-200) is evaluated as
        h1 = relu([e | upper] . [W_e | W_u]^T  +  cterm[frame of the row]),   cterm = c_frame . W_c^T + b,
    i.e. the conditioning block of the concat - constant over the FS samples of a frame - is multiplied once
    per (slot, frame) instead of once per sample: K drops from 3H to 2H in the three big comb
    GEMMs (forward, data gradient, weight gradient)."""

    @staticmethod
    def forward(ctx, xs_u8, conds, upper, target_u8, fused, emb, eg, ev, csw, csb, cw, cbias, w2g, w2v, b2, w3g, w3v, b3):
        dev = conds.device
        b, w = xs_u8.shape
        _, l, c = conds.shape
        h, q, r0 = ev.shape
        rf = w - r0 + 1
        m = b * rf
        fsz = rf // l
        cp = round_up(c, 8)
        assert q == 256, 'the fused log-softmax epilogue is built for q_levels == 256'

        # embedding + conv1d as a one-hot x table contraction: T[o, k*Q+q] = sum_q' We[o,q',k] E[q,q']
        onehot = ops.onehot_rows(xs_u8, q)                                   # (B, W, Q)
        e_b = ops.to_bf16(emb)                                               # E[q, q']
        we = _empty(h, r0 * q, device=dev)                                   # We[o, k*Q+q']
        we_t = _empty(q, r0 * h, device=dev)                                 # We^T[q', k*H+o]
        inv_e = _empty(h, dtype=F32, device=dev)
        ops.weight_prep(ev, eg, (h, q, r0), we, (r0 * q, 1, q), we_t, (1, r0 * h, h), inv_norm=inv_e)
        # tt[k*Q+q, o] = sum_q' E[q,q'] We[o,q',k]: the embedding + conv output for code q at tap k (the transposed table)
        tt = _empty(r0 * q, h, device=dev)
        for k in range(r0):
            ops.gemm_nt(e_b, we[:, k * q:], tt[k * q:], q, h, q, q, r0 * q, h)

        # comb_layer weights: W_e, W_u, W_c blocks (K-major) and the transposes for backward
        cwc = cw.contiguous()
        w_e = _empty(h, h, device=dev)
        w_c = _empty(h, h, device=dev)
        wcomb_t = _empty(3 * h, h, device=dev)                               # rows: e | c | upper blocks of W^T
        ops.weight_prep(cwc, None, (h, 3 * h, 1), wcomb_t, (1, h, 0))
        ops.weight_prep(cwc[:, :h].contiguous(), None, (h, h, 1), w_e, (h, 1, 0))
        ops.weight_prep(cwc[:, h:2 * h].contiguous(), None, (h, h, 1), w_c, (h, 1, 0))
        # The embedding, the conv AND comb_layer's embedding block are linear in the one-hot codes, so they fold into
        # ONE table T' = W_e . table (H x r0*Q); comb_layer's sample-rate input is then [one-hot windows | upper] and its
        # weight [T' | W_u]: the (B*RF, H) embedding activation, its GEMM and - in backward - its data and weight
        # gradient GEMMs over B*RF rows never exist.  (model.py:192-200)
        kc = r0 * q + h
        w_cat = _empty(h, kc, device=dev)
        ops.gemm_nt(w_e, tt, w_cat, h, r0 * q, h, h, h, kc)                   # T'[o', k*Q+q] = sum_o W_e[o',o] tt[k*Q+q,o]
        ops.weight_prep(cwc[:, 2 * h:].contiguous(), None, (h, h, 1), w_cat[:, r0 * q:], (kc, 1, 0))
        upper_c = upper.reshape(m, h)
        if not upper_c.is_contiguous() or upper_c.data_ptr() % 16:
            upper_c = upper_c.contiguous()

        # conditioning at frame rate: c_frame = conds_expand(conds) (model.py:194), cterm = c_frame W_c^T + b
        conds_b = _empty(b * l, cp, device=dev)
        ops.pad_cast_bf16(conds.contiguous(), b * l, c, c, conds_b, cp, cp)
        wcs = _zeros(h, cp, dtype=BF16, device=dev)
        wcs_t = _zeros(cp, h, dtype=BF16, device=dev)
        ops.weight_prep(csw.contiguous(), None, (h, c, 1), wcs, (cp, 1, 0), wcs_t, (1, h, 0))
        c_frame = _empty(b * l, h, device=dev)
        ops.gemm_nt(conds_b, wcs, c_frame, b * l, h, cp, cp, cp, h, bias=csb.contiguous())
        cterm = _empty(b * l, h, device=dev)
        ops.gemm_nt(c_frame, w_c, cterm, b * l, h, h, h, h, h, bias=cbias.contiguous())

        h1 = _empty(m, h, device=dev)
        # ReLU bit masks of h1 / h2 for the backward gates (4 bytes per row and 32 columns; see srnn_gemm_args.relu_mask)
        mk1 = torch.empty(m, (h + 31) // 32, dtype=torch.int32, device=dev)
        mk2 = torch.empty(m, (h + 31) // 32, dtype=torch.int32, device=dev)
        with ops.timed('comb_layer_fwd'):
            # A = [overlapping one-hot windows (row stride Q, row length r0*Q) | upper], batched per slot
            ops.gemm_nt(onehot, w_cat, h1, rf, h, kc, q, kc, h, batch=b, a_bs=w * q, c_bs=rf * h, aux=cterm, ldaux=h,
                        aux_bs=l * h, aux_mode=1, aux_row_div=fsz, relu=True, a2=upper_c, lda2=h, a2_bs=rf * h, k1=r0 * q,
                        relu_mask=mk1)
        w2 = _empty(h, h, device=dev)
        w2_t = _empty(h, h, device=dev)
        inv_2 = _empty(h, dtype=F32, device=dev)
        ops.weight_prep(w2v, w2g, (h, h, 1), w2, (h, 1, 0), w2_t, (1, h, 0), inv_norm=inv_2)
        h2 = _empty(m, h, device=dev)
        ops.gemm_nt(h1, w2, h2, m, h, h, h, h, h, bias=b2.contiguous(), relu=True, relu_mask=mk2)
        w3 = _empty(q, h, device=dev)
        w3_t = _empty(h, q, device=dev)
        inv_3 = _empty(q, dtype=F32, device=dev)
        ops.weight_prep(w3v, w3g, (q, h, 1), w3, (h, 1, 0), w3_t, (1, q, 0), inv_norm=inv_3)

        lse = _empty(m, dtype=F32, device=dev)
        logp_t = _empty(m, dtype=F32, device=dev)
        if target_u8 is None:
            target_u8 = torch.zeros(m, dtype=torch.uint8, device=dev)
        target_u8 = target_u8.contiguous()
        b3c = b3.contiguous()
        if fused:
            ops.gemm_nll(0, h2, w3, b3c, target_u8, m, h, h, h, lse=lse, logp_target=logp_t)
            out = logp_t.view(b, rf)
        else:
            logp = _empty(m, q, dtype=F32, device=dev)
            ops.gemm_nll(1, h2, w3, b3c, target_u8, m, h, h, h, lse=lse, logp_target=logp_t, logp=logp)
            out = logp.view(b, rf, q)
        ctx.dims = (b, w, l, c, h, q, r0, rf, m, fsz, cp, fused)
        ctx.save_for_backward(onehot, e_b, we_t, inv_e, conds_b, wcs_t, c_frame, upper_c, tt, wcomb_t, h1, w2, w2_t, inv_2,
                              h2, w3, w3_t, inv_3, target_u8, b3c, eg, ev, w2g, w2v, w3g, w3v, lse, mk1, mk2)
        return out

    @staticmethod
    def backward(ctx, gout):
        (onehot, e_b, we_t, inv_e, conds_b, wcs_t, c_frame, upper_c, tt, wcomb_t, h1, w2, w2_t, inv_2, h2, w3, w3_t, inv_3,
         target_u8, b3c, eg, ev, w2g, w2v, w3g, w3v, lse, mk1, mk2) = ctx.saved_tensors
        b, w, l, c, h, q, r0, rf, m, fsz, cp, fused = ctx.dims
        dev = gout.device
        gout = gout.contiguous().float()
        dlog = _empty(m, q, device=dev)
        if fused:
            ops.gemm_nll(2, h2, w3, b3c, target_u8, m, h, h, h, row_grad=gout, dlogits=dlog, lse=lse)
        else:
            ops.gemm_nll(3, h2, w3, b3c, target_u8, m, h, h, h, g=gout, dlogits=dlog, lse=lse)
        # adapt
        d_b3 = ops.colsum(dlog, m, q, q)
        dw3 = _zeros(q, h, device=dev)
        ops.gemm_tn(dlog, h2, dw3, q, h, m, q, h, h)
        d_w3v, d_w3g = ops.weight_prep_bwd(dw3, (h, 1, 0), w3v, w3g, inv_3, (q, h, 1))
        dh2 = _empty(m, h, device=dev)
        d_b2 = _zeros(h, dtype=F32, device=dev)                 # bias gradient = column sums, fused into the GEMM epilogue
        ops.gemm_nt(dlog, w3_t, dh2, m, h, q, q, q, h, gate_mask=mk2, colsum=d_b2)
        # comb_layer_expand
        dw2 = _zeros(h, h, device=dev)
        ops.gemm_tn(dh2, h1, dw2, h, h, m, h, h, h)
        d_w2v, d_w2g = ops.weight_prep_bwd(dw2, (h, 1, 0), w2v, w2g, inv_2, (h, h, 1))
        dh1 = _empty(m, h, device=dev)
        d_cbias = _zeros(h, dtype=F32, device=dev)
        ops.gemm_nt(dh2, w2_t, dh1, m, h, h, h, h, h, gate_mask=mk1, colsum=d_cbias)
        # comb_layer: [e | upper] blocks at sample rate, conditioning block at frame rate
        d_cw = _zeros(h, 3 * h, device=dev)
        ops.gemm_tn(dh1, upper_c, d_cw[:, 2 * h:], h, h, m, h, h, 3 * h)                   # d W_u
        seg = _empty(b * l, h, device=dev)                                   # sum of dh1 over the FS samples of a frame
        ops.repeat_rows_bwd(dh1, b * l, h, h, fsz, seg, h)
        ops.gemm_tn(seg, c_frame, d_cw[:, h:], h, h, b * l, h, h, 3 * h)                   # d W_c
        dupper = _empty(m, h, device=dev)
        ops.gemm_nt(dh1, wcomb_t[2 * h:], dupper, m, h, h, h, h, h)
        dc_frame = _empty(b * l, h, device=dev)
        ops.gemm_nt(seg, wcomb_t[h:], dc_frame, b * l, h, h, h, h, h)
        # conds_expand (frame rate)
        d_csb = ops.colsum(dc_frame, b * l, h, h)
        dwcs = _zeros(h, cp, device=dev)
        ops.gemm_tn(dc_frame, conds_b, dwcs, h, cp, b * l, h, cp, cp)
        dconds = _empty(b * l, c, dtype=F32, device=dev)
        ops.gemm_nt(dc_frame, wcs_t, dconds, b * l, c, h, h, h, c)
        # folded table: dT'[o', k*Q+q] = sum_{j: x[j+k]=q} dh1[j,o'].  T' = W_e . table, so d W_e = dT' . table^T and
        # d table = W_e^T . dT' are H x H x r0*Q contractions - no GEMM over the B*RF rows for the embedding side ...
        # ... as ONE TN GEMM over the overlapping windows: gt[k*Q+q, o'] = sum_j window_j[k*Q+q] dh1[j, o'] (A rows are
        # the r0*Q-wide windows at a row stride of Q, read in place)
        gt = _zeros(r0 * q, h, device=dev)
        ops.gemm_tn(onehot, dh1, gt, r0 * q, h, rf, q, h, h, batch=b, a_bs=w * q, b_bs=rf * h)
        # gt and G are fp32 sums over up to B*RF rows that feed further contractions: they enter those as hi + lo bf16
        # pairs (two accumulating passes / a K-concatenated operand), so no extra bf16 rounding sits in the chain
        gt_hi, gt_lo = ops.split_bf16(gt)
        for part in (gt_hi, gt_lo):
            ops.gemm_tn(part, tt, d_cw, h, h, r0 * q, h, h, 3 * h)           # d W_e[o',o] = sum_{kQ+q} gt[.,o'] tt[.,o]
        # G[q, k*H+o] = sum_o' gt[kQ+q,o'] W_e[o',o]: the gradient w.r.t. the (transposed) embedding + conv table
        g = _empty(q, r0 * h, dtype=F32, device=dev)
        gt2 = torch.cat((gt_hi, gt_lo), dim=1)                               # (r0*Q, 2H): [hi | lo] along K
        we_t2 = torch.cat((wcomb_t[:h], wcomb_t[:h]), dim=1)                 # (H, 2H): W_e^T twice
        for k in range(r0):
            ops.gemm_nt(gt2[k * q:], we_t2, g[:, k * h:], q, h, 2 * h, 2 * h, 2 * h, r0 * h)
        g_hi, g_lo = ops.split_bf16(g)
        d_emb = _empty(q, q, dtype=F32, device=dev)
        g2 = torch.cat((g_hi, g_lo), dim=1)                                  # (Q, 2*r0*H)
        we_t_2 = torch.cat((we_t, we_t), dim=1)
        ops.gemm_nt(g2, we_t_2, d_emb, q, q, 2 * r0 * h, 2 * r0 * h, 2 * r0 * h, q)
        dwe = _zeros(h, r0 * q, device=dev)
        for part in (g_hi, g_lo):
            for k in range(r0):
                ops.gemm_tn(part[:, k * h:], e_b, dwe[:, k * q:], h, q, q, r0 * h, q, r0 * q)
        d_ev, d_eg = ops.weight_prep_bwd(dwe, (r0 * q, 1, q), ev, eg, inv_e, (h, q, r0))
        return (None, dconds.view(b, l, c), dupper.view(b, rf, h), None, None,
                d_emb, d_eg.view_as(eg), d_ev, dwcs[:, :c].contiguous().view(h, c, 1), d_csb, d_cw, d_cbias,
                d_w2g.view_as(w2g), d_w2v, d_b2, d_w3g.view_as(w3g), d_w3v, d_b3)
