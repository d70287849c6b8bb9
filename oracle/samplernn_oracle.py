"""CPU oracle for the teacher-forced SampleRNN training step.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the arithmetic that
AlomdaElmasry/samplernn_pase performs on its hot path.  It exists so that the CUDA
path in ``samplernn_pase_b200`` can be checked on a GPU box where the reference
itself is not available.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against *outputs of the reference itself*, produced in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference``) and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.
The two extensions that have no reference implementation (LSTM tiers, PASE speaker
vector - BASELINE config 3) are marked "parity unpinned" at their definitions.

Style: everything is a pure function of an explicit parameter dict (keys = the
reference ``state_dict`` names) and explicit hidden state, i.e. the
"explicit-hidden functional core" underneath the reference's stateful modules.

Reference citations are ``file:line`` into the reference repository.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

MU = 255.0                       # utils.py:31
LOG_MU1 = 5.5451774444795623     # utils.py:32
EPS_LINEAR = 1e-2                # utils.py:29
EPS_ULAW = 1e-6                  # utils.py:30


# --------------------------------------------------------------------------------------
# Quantiser (utils.py:25-73)
# --------------------------------------------------------------------------------------
def quantize_ulaw(x: Tensor, q_levels: int = 256) -> Tensor:
    """utils.py:59-65.  Every step is a separate fp32 rounding, in this order.

    The final scale ``q_levels - 1e-6`` is a Python double that becomes 256.0f when it
    multiplies an fp32 tensor, so inputs >= ~0.9999997 map to index 256 (SURVEY trap 3).
    Runs on whatever device ``x`` lives on: on CUDA this is exactly the op chain the
    reference would launch (torch-CUDA multiplies by the fp32 reciprocal for the
    division; the CPU kernel divides - SURVEY A.1), which makes this function the
    bit-exact oracle for the CUDA quantiser when called with a CUDA tensor.
    """
    mag = (MU * x.abs() + 1.0).log()
    s = x.sign() * mag / LOG_MU1
    y = 0.5 * (s + 1.0)
    y = y * (q_levels - EPS_ULAW)
    return y.long()


def dequantize_ulaw(idx: Tensor, q_levels: int = 256) -> Tensor:
    """utils.py:67-73.  NB the sign is lost (``x`` is >= 0 before ``sign(x)*x``): the
    result is the magnitude only (SURVEY trap 2) and must be reproduced for parity."""
    y = idx.float() * 2.0 / q_levels - 1.0
    x = (y.abs() * LOG_MU1).exp() - 1
    return x.sign() * x / MU


def quantize_linear(x: Tensor, q_levels: int = 256) -> Tensor:
    """utils.py:48-54, restated with the per-row semantics the reference intends.

    The reference subtracts ``min(dim=-1)[0].expand_as(samples)`` which only broadcasts
    for 1-D inputs (or B == 1 / B == T); for those inputs this function is identical.
    """
    lo = x.min(dim=-1, keepdim=True)[0]
    y = x - lo
    hi = y.max(dim=-1, keepdim=True)[0]
    y = y / hi
    y = y * (q_levels - EPS_LINEAR)
    y = y + EPS_LINEAR / 2
    return y.long()


def dequantize_linear(idx: Tensor, q_levels: int = 256) -> Tensor:
    """utils.py:56-57."""
    return idx.float() / (q_levels / 2) - 1


def quantize(x: Tensor, ulaw: bool = True, q_levels: int = 256) -> Tensor:
    return quantize_ulaw(x, q_levels) if ulaw else quantize_linear(x, q_levels)  # utils.py:42-43


def dequantize(idx: Tensor, ulaw: bool = True, q_levels: int = 256) -> Tensor:
    return dequantize_ulaw(idx, q_levels) if ulaw else dequantize_linear(idx, q_levels)  # utils.py:45-46


def dequant_table(ulaw: bool = True, q_levels: int = 256) -> Tensor:
    """The 256(+1)-entry function table of ``dequantize``."""
    return dequantize(torch.arange(q_levels + 1), ulaw, q_levels)


# --------------------------------------------------------------------------------------
# Weight norm (torch.nn.utils.weight_norm, dim=0; model.py:135-138,183-186)
# --------------------------------------------------------------------------------------
def weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """w = g * v / ||v||, the norm taken over every dim except 0 (per slice of dim 0)."""
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)


# --------------------------------------------------------------------------------------
# Conditioning mixer (model.py:28-93)
# --------------------------------------------------------------------------------------
LING_CATEGORICAL = ([2, 3, 4, 5, 6], [27], [31, 33, 41], [49])          # model.py:80-85
LING_REAL_SLICES = ((0, 2), (7, 27), (28, 31), (32, 33), (34, 41), (42, 49), (50, None))  # model.py:86-92


def utterance_width(kind: str) -> int:
    return {'acoustic': 43, 'linguistic': 55, 'linguistic_lf0': 57}[kind]  # model.py:50-58


def expand_linguistic(p: Params, utt: Tensor, kind: str) -> Tensor:
    """model.py:76-93: 10 categorical columns -> embeddings, then the real columns."""
    if kind not in ('linguistic', 'linguistic_lf0'):
        return utt
    tables = ['conds_mixer.conds_utt_phonemes_emb.weight', 'conds_mixer.conds_utt_vowels_emb.weight',
              'conds_mixer.conds_utt_gpos_emb.weight', 'conds_mixer.conds_utt_tobi_emb.weight']
    parts = []
    for name, cols in zip(tables, LING_CATEGORICAL):
        for c in cols:
            parts.append(p[name][utt[:, :, c].long()])
    for a, b in LING_REAL_SLICES:
        parts.append(utt[:, :, a:b])
    return torch.cat(parts, dim=2)


def conds_mixer(p: Params, utt_conds: Tensor, speaker_ids: Optional[Tensor], kind: str = 'acoustic',
                speaker_vectors: Optional[Tensor] = None) -> Tensor:
    """model.py:60-65.  ``speaker_ids[i]`` is ``info[i]['speaker']['index']`` or 0 when the
    slot is empty (model.py:70).  Speaker block comes FIRST in the concat.

    ``speaker_vectors`` (B,S) is the O-D extension (PASE speaker vector): the reference
    raises for ``conds_speaker_type='pase'`` (model.py:73-74) - PARITY UNPINNED; the vector
    simply replaces the embedding row.
    """
    b, l, _ = utt_conds.shape
    if speaker_vectors is None:
        spk = p['conds_mixer.speaker_embedding.weight'][speaker_ids]      # model.py:67-72
    else:
        spk = speaker_vectors
    spk = spk.unsqueeze(1).expand(b, l, -1)
    feats = torch.cat((spk, expand_linguistic(p, utt_conds, kind)), dim=2)
    return feats @ p['conds_mixer.conds_mix.weight'].t() + p['conds_mixer.conds_mix.bias']


# --------------------------------------------------------------------------------------
# Recurrent cells
# --------------------------------------------------------------------------------------
def gru_sequence(x: Tensor, h0: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor
                 ) -> Tuple[Tensor, Tensor]:
    """One GRU layer over time, written out (torch.nn.GRU semantics, model.py:110,152):
    gates stacked (r,z,n); n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1-z) n + z h."""
    hsz = h0.shape[1]
    gi_all = x @ w_ih.t() + b_ih
    h = h0
    outs = []
    for t in range(x.shape[1]):
        gi = gi_all[:, t]
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :hsz] + gh[:, :hsz])
        z = torch.sigmoid(gi[:, hsz:2 * hsz] + gh[:, hsz:2 * hsz])
        n = torch.tanh(gi[:, 2 * hsz:] + r * gh[:, 2 * hsz:])
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, dim=1), h


def lstm_sequence(x: Tensor, h0: Tensor, c0: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """O-C extension, PARITY UNPINNED (the reference has GRU only, model.py:110).
    torch.nn.LSTM semantics: gates stacked (i,f,g,o)."""
    hsz = h0.shape[1]
    gi_all = x @ w_ih.t() + b_ih
    h, c = h0, c0
    outs = []
    for t in range(x.shape[1]):
        g = gi_all[:, t] + h @ w_hh.t() + b_hh
        i = torch.sigmoid(g[:, :hsz])
        f = torch.sigmoid(g[:, hsz:2 * hsz])
        gg = torch.tanh(g[:, 2 * hsz:3 * hsz])
        o = torch.sigmoid(g[:, 3 * hsz:])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1), h, c


def _fast_gru(x, h0_layers, weights):
    """Library-fused equivalent of stacking ``gru_sequence`` (used only for timing the
    CPU baseline; tests check it equals the written-out loop)."""
    flat = [w for layer in weights for w in layer]
    # (train=True only selects cuDNN's training workspace - its backward refuses to run otherwise; dropout is 0)
    out, hn = torch._VF.gru(x, h0_layers, flat, True, len(weights), 0.0, True, False, True)
    return out, hn


# --------------------------------------------------------------------------------------
# Frame-level tier (model.py:96-156)
# --------------------------------------------------------------------------------------
def frame_tier(p: Params, n: int, frames: Tensor, conds: Tensor, upper: Optional[Tensor], h0: Tensor,
               cell: str = 'gru', c0: Optional[Tensor] = None, fast: bool = False):
    """``frames`` (B,T,fs_n) dequantised samples, ``conds`` (B,L,C), ``upper`` (B,T,H) or None,
    ``h0`` (layers,B,H).  Returns (upsampled (B,T*r,H), h_n (layers,B,H)[, c_n])."""
    pre = f'frames_layers.{n}.'
    b, t, _ = frames.shape
    rep = t // conds.shape[1]
    c = conds.repeat_interleave(rep, dim=1) if rep != 1 else conds                  # model.py:142-145
    lib = fast == 'library'          # the operators the reference itself launches (Conv1d / ConvTranspose1d -> cuDNN / oneDNN)
    if lib:
        wx3 = weight_norm(p[pre + 'x_expand.weight_g'], p[pre + 'x_expand.weight_v'])
        wc3 = weight_norm(p[pre + 'conds_expand.weight_g'], p[pre + 'conds_expand.weight_v'])
        u = F.conv1d(frames.permute(0, 2, 1), wx3, p[pre + 'x_expand.bias']).permute(0, 2, 1) + \
            F.conv1d(c.permute(0, 2, 1), wc3, p[pre + 'conds_expand.bias']).permute(0, 2, 1)      # model.py:146-147
    else:
        wx = weight_norm(p[pre + 'x_expand.weight_g'], p[pre + 'x_expand.weight_v'])[:, :, 0]
        wc = weight_norm(p[pre + 'conds_expand.weight_g'], p[pre + 'conds_expand.weight_v'])[:, :, 0]
        u = frames @ wx.t() + p[pre + 'x_expand.bias'] + c @ wc.t() + p[pre + 'conds_expand.bias']  # :146-147
    if upper is not None:
        u = u + upper                                                                # model.py:148
    layers = h0.shape[0]
    hn, cn = [], []
    if cell == 'gru' and fast:
        ws = [[p[pre + f'rnn.weight_ih_l{l}'], p[pre + f'rnn.weight_hh_l{l}'],
               p[pre + f'rnn.bias_ih_l{l}'], p[pre + f'rnn.bias_hh_l{l}']] for l in range(layers)]
        u, hn_t = _fast_gru(u, h0.contiguous(), ws)
        hn = list(hn_t)
    else:
        for l in range(layers):                                                      # model.py:152
            args = (p[pre + f'rnn.weight_ih_l{l}'], p[pre + f'rnn.weight_hh_l{l}'],
                    p[pre + f'rnn.bias_ih_l{l}'], p[pre + f'rnn.bias_hh_l{l}'])
            if cell == 'gru':
                u, h_last = gru_sequence(u, h0[l], *args)
            else:
                u, h_last, c_last = lstm_sequence(u, h0[l], c0[l], *args)
                cn.append(c_last)
            hn.append(h_last)
    # learned upsampling: ConvTranspose1d(H,H,r,stride=r) + upsample_bias(H,r)   model.py:153-155
    wu = weight_norm(p[pre + 'upsample.weight_g'], p[pre + 'upsample.weight_v'])     # (H_in, H_out, r)
    r = wu.shape[2]
    if lib:
        bias = p[pre + 'upsample_bias'].unsqueeze(0).unsqueeze(2).expand(b, wu.shape[1], t, r).contiguous() \
            .view(b, wu.shape[1], t * r)                                             # model.py:153-154
        up = (F.conv_transpose1d(u.permute(0, 2, 1), wu, stride=r) + bias).permute(0, 2, 1)   # model.py:155
    else:
        up = torch.einsum('bti,ioj->btjo', u, wu) + p[pre + 'upsample_bias'].t()     # (B,T,r,H)
        up = up.reshape(b, t * r, -1)
    if cell == 'gru':
        return up, torch.stack(hn)
    return up, torch.stack(hn), torch.stack(cn)


# --------------------------------------------------------------------------------------
# Sample-level MLP (model.py:159-203)
# --------------------------------------------------------------------------------------
def sample_level(p: Params, xs: Tensor, conds: Tensor, upper: Tensor, fast=False) -> Tensor:
    """``xs`` (B,RF+r0-1) int64, ``conds`` (B,L,C), ``upper`` (B,RF,H) -> log-probs (B,RF,Q).
    ``fast == 'library'``: the same arithmetic through the operators the reference launches (Conv1d for the
    embedding conv and every 1x1 projection, Linear for comb_layer; model.py:188-203) - used for timing."""
    pre = 'sample_layer.'
    b, rf, h = upper.shape
    if fast == 'library':
        emb = p[pre + 'emb_layer.weight'][xs.reshape(-1)].view(b, -1, p[pre + 'emb_layer.weight'].shape[1])
        we = weight_norm(p[pre + 'emb_layer_expand.weight_g'], p[pre + 'emb_layer_expand.weight_v'])
        e = F.conv1d(emb.permute(0, 2, 1), we)                                       # model.py:193  (B,H,RF)
        rep = rf // conds.shape[1]
        c = conds.unsqueeze(2).expand(b, conds.shape[1], rep, conds.shape[2]).reshape(b, rf, conds.shape[2])
        c = F.conv1d(c.permute(0, 2, 1), p[pre + 'conds_expand.weight'], p[pre + 'conds_expand.bias'])
        cat = torch.cat((e.permute(0, 2, 1), c.permute(0, 2, 1), upper), dim=2)
        h1 = F.relu(F.linear(cat, p[pre + 'comb_layer.weight'], p[pre + 'comb_layer.bias']))
        w2 = weight_norm(p[pre + 'comb_layer_expand.weight_g'], p[pre + 'comb_layer_expand.weight_v'])
        h2 = F.relu(F.conv1d(h1.permute(0, 2, 1), w2, p[pre + 'comb_layer_expand.bias']))
        w3 = weight_norm(p[pre + 'adapt.weight_g'], p[pre + 'adapt.weight_v'])
        logits = F.conv1d(h2, w3, p[pre + 'adapt.bias'])
        return F.log_softmax(logits.permute(0, 2, 1), dim=2)
    emb = p[pre + 'emb_layer.weight'][xs]                                            # model.py:192
    we = weight_norm(p[pre + 'emb_layer_expand.weight_g'], p[pre + 'emb_layer_expand.weight_v'])  # (H,Q,r0)
    r0 = we.shape[2]
    win = emb.unfold(1, r0, 1)                                                       # (B,RF,Q,r0)
    e = torch.einsum('bjqk,oqk->bjo', win, we)                                       # model.py:193
    c = conds.repeat_interleave(rf // conds.shape[1], dim=1)                         # model.py:189-191
    cw = p[pre + 'conds_expand.weight'][:, :, 0]
    c = c @ cw.t() + p[pre + 'conds_expand.bias']                                    # model.py:194
    cat = torch.cat((e, c, upper), dim=2)                                            # model.py:196-199
    h1 = F.relu(cat @ p[pre + 'comb_layer.weight'].t() + p[pre + 'comb_layer.bias'])  # model.py:195
    w2 = weight_norm(p[pre + 'comb_layer_expand.weight_g'], p[pre + 'comb_layer_expand.weight_v'])[:, :, 0]
    h2 = F.relu(h1 @ w2.t() + p[pre + 'comb_layer_expand.bias'])                     # model.py:201
    w3 = weight_norm(p[pre + 'adapt.weight_g'], p[pre + 'adapt.weight_v'])[:, :, 0]
    logits = h2 @ w3.t() + p[pre + 'adapt.bias']                                     # model.py:202
    return F.log_softmax(logits, dim=2)                                              # model.py:203


# --------------------------------------------------------------------------------------
# Whole model (model.py:206-287)
# --------------------------------------------------------------------------------------
class ModelSpec:
    """The constructor arguments of ``SampleRNNModel`` (model.py:212-214) that matter."""

    def __init__(self, ratios: Sequence[int], rnn_layers: Sequence[int], rnn_hidden_size: Sequence[int],
                 sequence_length: int, conds_utterance_type: str = 'acoustic', q_type_ulaw: bool = True,
                 q_levels: int = 256, cell: str = 'gru'):
        self.ratios = list(ratios)
        self.rnn_layers = list(rnn_layers)
        self.hidden = list(rnn_hidden_size)
        self.sequence_length = sequence_length
        self.kind = conds_utterance_type
        self.ulaw = q_type_ulaw
        self.q_levels = q_levels
        self.cell = cell
        self.frame_sizes = [int(math.prod(self.ratios[:i + 1])) for i in range(len(self.ratios))]  # model.py:226
        self.frame_size = self.frame_sizes[-1]                                       # model.py:215
        self.receptive_field = self.frame_size * sequence_length                     # model.py:216


def initial_state(p: Params, spec: ModelSpec, n: int, reset: Sequence[int], stored: Optional[Tensor],
                  stored_valid: Optional[Sequence[bool]], key: str = 'rnn_h0') -> Tensor:
    """model.py:149-151 + 239-243, dense form: slot i starts from its carried
    state iff reset[i]==0 and a state is stored, else from the learnable ``rnn_h0``."""
    h0 = p[f'frames_layers.{n}.{key}']
    cols = []
    for i, r in enumerate(reset):
        if int(r) == 0 and stored is not None and stored_valid[i]:
            cols.append(stored[:, i])
        else:
            cols.append(h0)
    return torch.stack(cols, dim=1)


class CarryState:
    """Per-tier carried hidden state: tensors (layers,B,H) + per-slot validity."""

    def __init__(self):
        self.h: Dict[int, Tensor] = {}
        self.c: Dict[int, Tensor] = {}
        self.valid: Dict[int, List[bool]] = {}


def forward(p: Params, spec: ModelSpec, x: Tensor, y: Tensor, utt_conds: Tensor, speaker_ids: Tensor,
            reset: Sequence[int], state: Optional[CarryState] = None, carry: bool = True,
            speaker_vectors: Optional[Tensor] = None, fast: bool = False):
    """model.py:252-287.  ``carry=False`` reproduces the reference *as written* (O-A: the
    ``hasattr(self,'rnnstates')`` typo at model.py:256 re-creates the state store on every
    call so every chunk starts from ``rnn_h0``); ``carry=True`` is the documented intent
    (O-B, README.md:17-21).  Returns (log-probs of valid slots, targets of valid slots,
    new CarryState, dict of intermediates)."""
    xq = quantize(x, spec.ulaw, spec.q_levels)                                       # model.py:260
    yq = quantize(y, spec.ulaw, spec.q_levels)
    return forward_indices(p, spec, xq, yq, utt_conds, speaker_ids, reset, state, carry, speaker_vectors, fast)


def forward_indices(p: Params, spec: ModelSpec, xq: Tensor, yq: Tensor, utt_conds: Tensor, speaker_ids: Tensor,
                    reset: Sequence[int], state: Optional[CarryState] = None, carry: bool = True,
                    speaker_vectors: Optional[Tensor] = None, fast: bool = False):
    """``forward`` after the quantiser (model.py:263-287): takes the int64 indices directly.  Used to
    teacher-force a *generated* index sequence (SURVEY probe P8)."""
    reset = [int(r) for r in reset]
    if state is None or not carry:
        state = CarryState()
    new_state = CarryState()
    x, y = xq, yq
    conds = conds_mixer(p, utt_conds, speaker_ids, spec.kind, speaker_vectors)       # model.py:263
    fs_top = spec.frame_size
    rf = y.shape[1]
    upper = None
    for n in reversed(range(len(spec.ratios))):                                      # model.py:267
        fs = spec.frame_sizes[n]
        sl = xq[:, fs_top - fs: fs_top - fs + rf]                                    # model.py:268-270 (A.2)
        frames = dequantize(sl, spec.ulaw, spec.q_levels).reshape(x.shape[0], -1, fs)  # model.py:271
        h0 = initial_state(p, spec, n, reset, state.h.get(n), state.valid.get(n))
        if spec.cell == 'gru':
            upper, hn = frame_tier(p, n, frames, conds, upper, h0, fast=fast)        # model.py:273
        else:
            c0 = initial_state(p, spec, n, reset, state.c.get(n), state.valid.get(n), key='rnn_c0')
            upper, hn, cn = frame_tier(p, n, frames, conds, upper, h0, cell='lstm', c0=c0)
            new_state.c[n] = cn.detach()
        new_state.h[n] = hn.detach()                                                 # model.py:276
        new_state.valid[n] = [r in (0, 1) for r in reset]                            # model.py:245-250
    xs = xq[:, fs_top - spec.ratios[0]:]                                             # model.py:279
    logp = sample_level(p, xs, conds, upper, fast=fast)                              # model.py:280
    keep = torch.tensor([r != 2 for r in reset])
    return logp[keep], yq[keep], new_state, {'xq': xq, 'yq': yq, 'conds': conds, 'upper': upper}  # :283-284


def nll(logp: Tensor, target: Tensor) -> Tensor:
    """runner.py:52."""
    return F.nll_loss(logp.reshape(-1, logp.shape[2]), target.reshape(-1))


# --------------------------------------------------------------------------------------
# Optimiser (optimizer.py:6-14)
# --------------------------------------------------------------------------------------
def adam_clipped_step(params: List[Tensor], grads: List[Tensor], exp_avg: List[Tensor], exp_avg_sq: List[Tensor],
                      step: int, lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """Clamp every gradient to [-1,1] (optimizer.py:12) then a plain Adam update
    (torch.optim.Adam defaults, no weight decay, no amsgrad).  ``step`` is 1-based."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    for w, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g.clamp(-1.0, 1.0)
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        w.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# Parameter construction (same distributions as model.py:8-25,118-133,176-181; used by the
# CPU baseline and by tests that do not have reference weights at hand)
# --------------------------------------------------------------------------------------
def _uniform(shape, bound, gen):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def init_params(spec: ModelSpec, conds_speaker_n: int = 126, conds_speaker_size: int = 15, conds_size: int = 50,
                ling_n: Sequence[int] = (64, 32, 16, 8), ling_emb: int = 10, seed: int = 1234,
                perturb: float = 0.0) -> Params:
    """Random parameters with the reference's shapes/key names.  Distributions
    follow the reference initialisers in spirit (kaiming/lecun uniform, orthogonal n-gate);
    parity tests that need *identical* weights load golden state_dicts instead.
    ``perturb`` adds N(0,perturb^2) to the zero-initialised tensors so they cannot hide bugs."""
    g = torch.Generator().manual_seed(seed)
    q = spec.q_levels
    p: Params = {}
    u_raw = utterance_width(spec.kind)
    u_exp = u_raw if spec.kind == 'acoustic' else u_raw - 10 + 10 * ling_emb
    p['conds_mixer.speaker_embedding.weight'] = torch.randn(conds_speaker_n, conds_speaker_size, generator=g)
    if spec.kind != 'acoustic':
        for name, nn_ in zip(('phonemes', 'vowels', 'gpos', 'tobi'), ling_n):
            p[f'conds_mixer.conds_utt_{name}_emb.weight'] = torch.randn(nn_, ling_emb, generator=g)
    fan = u_exp + conds_speaker_size
    p['conds_mixer.conds_mix.weight'] = _uniform((conds_size, fan), 1 / math.sqrt(fan), g)
    p['conds_mixer.conds_mix.bias'] = _uniform((conds_size,), 1 / math.sqrt(fan), g)
    gates = 3 if spec.cell == 'gru' else 4

    def zeros(*shape):
        t = torch.zeros(*shape)
        return t + perturb * torch.randn(*shape, generator=g) if perturb else t

    def normed(prefix, shape, bound):
        v = _uniform(shape, bound, g)
        p[prefix + '.weight_v'] = v
        gshape = (shape[0],) + (1,) * (len(shape) - 1)
        p[prefix + '.weight_g'] = v.reshape(shape[0], -1).norm(dim=1).reshape(gshape).clone()

    for n, (fs, r, layers, h) in enumerate(zip(spec.frame_sizes, spec.ratios, spec.rnn_layers, spec.hidden)):
        pre = f'frames_layers.{n}'
        p[pre + '.rnn_h0'] = zeros(layers, h)
        if spec.cell == 'lstm':
            p[pre + '.rnn_c0'] = zeros(layers, h)
        p[pre + '.upsample_bias'] = zeros(h, r)
        normed(pre + '.x_expand', (h, fs, 1), math.sqrt(6 / fs))
        p[pre + '.x_expand.bias'] = zeros(h)
        normed(pre + '.conds_expand', (h, conds_size, 1), math.sqrt(6 / conds_size))
        p[pre + '.conds_expand.bias'] = zeros(h)
        for l in range(layers):
            p[pre + f'.rnn.weight_ih_l{l}'] = _uniform((gates * h, h), math.sqrt(3 / h), g)
            whh = _uniform((gates * h, h), math.sqrt(3 / h), g)
            qmat, _ = torch.linalg.qr(torch.randn(h, h, generator=g))
            whh[(gates - 2 if spec.cell == 'lstm' else gates - 1) * h:][:h] = qmat
            p[pre + f'.rnn.weight_hh_l{l}'] = whh
            p[pre + f'.rnn.bias_ih_l{l}'] = zeros(gates * h)
            p[pre + f'.rnn.bias_hh_l{l}'] = zeros(gates * h)
        normed(pre + '.upsample', (h, h, r), math.sqrt(6 / h))
    h = spec.hidden[0]
    r0 = spec.ratios[0]
    pre = 'sample_layer'
    p[pre + '.emb_layer.weight'] = torch.randn(q, q, generator=g)
    normed(pre + '.emb_layer_expand', (h, q, r0), math.sqrt(6 / (q * r0)))
    p[pre + '.conds_expand.weight'] = _uniform((h, conds_size, 1), 1 / math.sqrt(conds_size), g)
    p[pre + '.conds_expand.bias'] = _uniform((h,), 1 / math.sqrt(conds_size), g)
    p[pre + '.comb_layer.weight'] = _uniform((h, 3 * h), math.sqrt(6 / (3 * h)), g)
    p[pre + '.comb_layer.bias'] = zeros(h)
    normed(pre + '.comb_layer_expand', (h, h, 1), 1 / math.sqrt(h))
    p[pre + '.comb_layer_expand.bias'] = _uniform((h,), 1 / math.sqrt(h), g)
    normed(pre + '.adapt', (q, h, 1), math.sqrt(3 / h))
    p[pre + '.adapt.bias'] = zeros(q)
    return p


# --------------------------------------------------------------------------------------
# Synthetic workload (SURVEY 8(d) "Synthetic inputs") and a whole train step
# --------------------------------------------------------------------------------------
def synthetic_utterances(spec: ModelSpec, batch: int, chunks: int, seed: int = 4321, conds_width: Optional[int] = None,
                         n_speakers: int = 126):
    """wav ~ U(-0.99,0.99) with FS leading zeros (dataset.py:50); conds ~ N(0,1); speaker i % n."""
    g = torch.Generator().manual_seed(seed)
    fs, rf, l = spec.frame_size, spec.receptive_field, spec.sequence_length
    width = conds_width or utterance_width(spec.kind)
    wav = (torch.rand(batch, fs + chunks * rf, generator=g) * 2 - 1) * 0.99
    wav[:, :fs] = 0.0
    conds = torch.randn(batch, chunks * l, width, generator=g)
    if spec.kind != 'acoustic':
        for cols, ncat in zip(LING_CATEGORICAL, (64, 32, 16, 8)):
            for c in cols:
                conds[:, :, c] = torch.randint(0, ncat, (batch, chunks * l), generator=g).float()
    speakers = torch.arange(batch) % n_speakers
    return wav, conds, speakers


def chunk_of(spec: ModelSpec, wav: Tensor, conds: Tensor, k: int):
    """loader.py:76-77,83-84: x = wav[k*RF : k*RF+RF+FS-1], y = wav[FS+k*RF : FS+(k+1)*RF]."""
    fs, rf, l = spec.frame_size, spec.receptive_field, spec.sequence_length
    x = wav[:, k * rf: k * rf + rf + fs - 1]
    y = wav[:, fs + k * rf: fs + (k + 1) * rf]
    return x.contiguous(), y.contiguous(), conds[:, k * l:(k + 1) * l].contiguous()


class CpuTrainer:
    """forward + NLL + backward + AdamClipped with plain torch ops on whatever device the parameters live on:
    the ``cpu_baseline`` / ``--impl reference`` timing (CPU) and the informational GPU-eager leg (cuda) of bench.py.
    ``fast='library'`` runs the reference's own operator choice (fused GRU, Conv1d, ConvTranspose1d, Linear)."""

    def __init__(self, spec: ModelSpec, params: Params, lr: float = 1e-4, fast='library'):
        self.spec = spec
        self.fast = fast if spec.cell == 'gru' else True
        self.params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.m = [torch.zeros_like(v) for v in self.params.values()]
        self.v = [torch.zeros_like(v) for v in self.params.values()]
        self.t = 0
        self.lr = lr
        self.state: Optional[CarryState] = None

    def step(self, x, y, conds, speakers, reset) -> float:
        for v in self.params.values():
            v.grad = None
        logp, tgt, self.state, _ = forward(self.params, self.spec, x, y, conds, speakers, reset, self.state,
                                            carry=True, fast=self.fast)
        loss = nll(logp, tgt)
        loss.backward()
        self.t += 1
        with torch.no_grad():
            ws = list(self.params.values())
            adam_clipped_step(ws, [w.grad if w.grad is not None else torch.zeros_like(w) for w in ws],
                              self.m, self.v, self.t, self.lr)
        return float(loss)
