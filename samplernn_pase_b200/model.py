"""B200-native SampleRNN modules behind the reference's ``model.py`` API.

Same class names, constructor arguments, attributes, ``forward``/``test`` signatures and
``state_dict`` keys as ``samplernn_pase/model.py`` so the Skeltorch runner
(runner.py:11-27,47-52) can use this module unchanged.  The modules only *hold* fp32 parameters;
all arithmetic runs in the sm_100a kernels of ``csrc/`` through ``functional.py``.

Differences a caller can observe (all documented in DESIGN.md):
  * arithmetic is bf16 x bf16 -> fp32 (tensor cores) instead of fp32; ``precision='fp32'`` selects the fp32-tolerance
    mode of ``functional_f32.py`` (split-bf16 operands, fp32 activations; ~4x slower, GRU tiers only, training
    forward/backward only - generation always runs in bf16);
  * hidden-state carry-over works by default (the reference's ``hasattr(self, 'rnnstates')``
    typo, model.py:256, disables it); ``reference_as_written=True`` reproduces the typo;
  * ``fused_loss=True`` makes ``forward`` return ``(log p(target) as (B_valid, RF, 1), zeros)`` so
    that the runner's ``F.nll_loss(y_hat.view(-1, y_hat.size(2)), y.view(-1))`` (runner.py:52)
    yields the same scalar without materialising the (B, RF, 256) log-probabilities.
"""
import math

import numpy as np
import torch

from . import ops
from .functional import CondsMixFn, FrameTierFn, SampleLevelFn, StateSelectFn
from .functional_f32 import CondsMixFn32, FrameTierFn32, SampleLevelFn32
from .utils import SampleRNNQuantizer


def _uniform_(t, bound):
    return torch.nn.init.uniform_(t, -bound, bound)


def _lecun_uniform_(t):
    fan_in = t.shape[1] * (t[0][0].numel() if t.dim() > 2 else 1)
    return _uniform_(t, math.sqrt(3.0 / fan_in))                               # model.py:8-10


class _Params(torch.nn.Module):
    """A bag of named parameters (keeps the reference's dotted state_dict names)."""

    def __init__(self, **tensors):
        super().__init__()
        for name, t in tensors.items():
            self.register_parameter(name, torch.nn.Parameter(t))


def _normed(v, bias=None):
    """weight-norm parametrisation of ``v`` over dim 0: g initialised to ||v|| (model.py:135-138)."""
    g = v.reshape(v.shape[0], -1).norm(dim=1).reshape((v.shape[0],) + (1,) * (v.dim() - 1))
    fields = {}
    if bias is not None:
        fields['bias'] = bias
    fields['weight_g'] = g
    fields['weight_v'] = v
    return _Params(**fields)


class CondsMixer(torch.nn.Module):
    """model.py:28-93."""

    def __init__(self, conds_speaker_type, conds_speaker_n, conds_speaker_size, conds_utterance_type,
                 conds_utterance_linguistic_n, conds_utterance_linguistic_emb_size, conds_size):
        super().__init__()
        self.conds_speaker_type = conds_speaker_type
        self.conds_speaker_size = conds_speaker_size
        self.conds_utterance_type = conds_utterance_type
        emb = conds_utterance_linguistic_emb_size
        sizes = {'acoustic': (43, 43), 'linguistic': (55, 55 - 10 + 10 * emb), 'linguistic_lf0': (57, 57 - 10 + 10 * emb)}
        self.conds_utterance_size, self.conds_utterance_expanded_size = sizes[conds_utterance_type]   # model.py:50-58
        self.speaker_embedding = _Params(weight=torch.randn(conds_speaker_n, conds_speaker_size))
        if conds_utterance_type in ('linguistic', 'linguistic_lf0'):
            n = conds_utterance_linguistic_n
            self.conds_utt_phonemes_emb = _Params(weight=torch.randn(n[0], emb))
            self.conds_utt_vowels_emb = _Params(weight=torch.randn(n[1], emb))
            self.conds_utt_gpos_emb = _Params(weight=torch.randn(n[2], emb))
            self.conds_utt_tobi_emb = _Params(weight=torch.randn(n[3], emb))
        fan_in = self.conds_utterance_expanded_size + conds_speaker_size
        bound = 1.0 / math.sqrt(fan_in)
        self.conds_mix = _Params(weight=_uniform_(torch.empty(conds_size, fan_in), bound),
                                 bias=_uniform_(torch.empty(conds_size), bound))

    def speaker_ids(self, info, device):
        """model.py:67-72: empty slots (info None) use speaker 0."""
        ids = [it['speaker']['index'] if it is not None else 0 for it in info]
        return torch.tensor(ids, dtype=torch.int32).to(device, non_blocking=True)

    def speaker_vectors(self, info, device):
        """``conds_speaker_type='pase'``: the reference raises (model.py:73-74); this extension (BASELINE
        config 3, oracle variant O-D, parity unpinned) takes a pre-computed PASE speaker vector of width
        ``conds_speaker_size`` from ``info[i]['speaker']['pase']`` in place of the embedding row (zeros for
        an empty slot)."""
        rows = [torch.as_tensor(it['speaker']['pase'], dtype=torch.float32) if it is not None
                else torch.zeros(self.conds_speaker_size) for it in info]
        return torch.stack(rows).to(device, non_blocking=True)

    def _expand_linguistic(self, utt):
        """model.py:76-93.  Index gathers + concat (data movement; the tables' gradients flow
        back through ``CondsMixFn``'s d(utt))."""
        if self.conds_utterance_type not in ('linguistic', 'linguistic_lf0'):
            return utt
        parts = [self.conds_utt_phonemes_emb.weight[utt[:, :, i].long()] for i in (2, 3, 4, 5, 6)]
        parts.append(self.conds_utt_vowels_emb.weight[utt[:, :, 27].long()])
        parts += [self.conds_utt_gpos_emb.weight[utt[:, :, i].long()] for i in (31, 33, 41)]
        parts.append(self.conds_utt_tobi_emb.weight[utt[:, :, 49].long()])
        for a, b in ((0, 2), (7, 27), (28, 31), (32, 33), (34, 41), (42, 49), (50, None)):
            parts.append(utt[:, :, a:b])
        return torch.cat(parts, dim=2)

    def forward(self, utt_conds, info):
        if self.conds_speaker_type == 'pase':
            table = self.speaker_vectors(info, utt_conds.device)               # one row per slot
            ids = torch.arange(len(info), dtype=torch.int32, device=utt_conds.device)
        else:
            table = self.speaker_embedding.weight
            ids = self.speaker_ids(info, utt_conds.device)
        fn = CondsMixFn32 if getattr(self, 'precision', 'bf16') == 'fp32' else CondsMixFn
        return fn.apply(self._expand_linguistic(utt_conds), ids, table, self.conds_mix.weight, self.conds_mix.bias)


class _RnnParams(torch.nn.Module):
    def __init__(self, hidden, layers, gates=3):
        super().__init__()
        for l in range(layers):
            w_ih = torch.empty(gates * hidden, hidden)
            w_hh = torch.empty(gates * hidden, hidden)
            for g in range(gates):                                             # model.py:129-132 (per-gate init)
                _uniform_(w_ih[g * hidden:(g + 1) * hidden], math.sqrt(3.0 / hidden))
                if g == gates - 1:
                    torch.nn.init.orthogonal_(w_hh[g * hidden:(g + 1) * hidden])
                else:
                    _uniform_(w_hh[g * hidden:(g + 1) * hidden], math.sqrt(3.0 / hidden))
            self.register_parameter(f'weight_ih_l{l}', torch.nn.Parameter(w_ih))
            self.register_parameter(f'weight_hh_l{l}', torch.nn.Parameter(w_hh))
            self.register_parameter(f'bias_ih_l{l}', torch.nn.Parameter(torch.zeros(gates * hidden)))
            self.register_parameter(f'bias_hh_l{l}', torch.nn.Parameter(torch.zeros(gates * hidden)))

    def flat(self, layers):
        out = []
        for l in range(layers):
            out += [getattr(self, f'weight_ih_l{l}'), getattr(self, f'weight_hh_l{l}'),
                    getattr(self, f'bias_ih_l{l}'), getattr(self, f'bias_hh_l{l}')]
        return out


class FrameLevelLayer(torch.nn.Module):
    """model.py:96-156."""

    def __init__(self, input_samples, conds_size, ratio, rnn_layers, rnn_hidden_size, rnn_cell='gru'):
        super().__init__()
        self.input_samples = input_samples
        self.ratio = ratio
        self.rnn_layers = rnn_layers
        self.rnn_hidden_size = rnn_hidden_size
        self.rnn_cell = rnn_cell                      # 'lstm' is an extension (the reference has GRU only)
        h = rnn_hidden_size
        self.rnn_h0 = torch.nn.Parameter(torch.zeros(rnn_layers, h))
        if rnn_cell == 'lstm':
            self.rnn_c0 = torch.nn.Parameter(torch.zeros(rnn_layers, h))
        self.upsample_bias = torch.nn.Parameter(torch.zeros(h, ratio))
        kaiming = lambda shape, fan_in: _uniform_(torch.empty(shape), math.sqrt(6.0 / fan_in))   # noqa: E731
        self.x_expand = _normed(kaiming((h, input_samples, 1), input_samples), torch.zeros(h))   # model.py:119,121
        self.conds_expand = _normed(kaiming((h, conds_size, 1), conds_size), torch.zeros(h))     # model.py:120,122
        self.rnn = _RnnParams(h, rnn_layers, gates=4 if rnn_cell == 'lstm' else 3)
        self.upsample = _normed(_uniform_(torch.empty(h, h, ratio), math.sqrt(6.0 / h)))          # model.py:124-126

    def _tier(self, xq_u8, x_off, lut, frames, conds, upper, h_init, into_cat=False, c_init=None):
        """-> (upsampled, h_n, c_n); c_n is an empty tensor for GRU tiers."""
        fn = FrameTierFn32 if getattr(self, 'precision', 'bf16') == 'fp32' else FrameTierFn
        return fn.apply(
            xq_u8, x_off, lut, frames, conds, upper, h_init, self.input_samples, self.ratio, into_cat, c_init,
            self.x_expand.weight_g, self.x_expand.weight_v, self.x_expand.bias,
            self.conds_expand.weight_g, self.conds_expand.weight_v, self.conds_expand.bias,
            self.upsample.weight_g, self.upsample.weight_v, self.upsample_bias, *self.rnn.flat(self.rnn_layers))

    def initial_state(self, carried, use_carry):
        return StateSelectFn.apply(self.rnn_h0, carried, use_carry)

    def initial_cell(self, carried, use_carry):
        return StateSelectFn.apply(self.rnn_c0, carried, use_carry)

    def forward(self, x, conds, upper_conditioning, rnn_state):
        """Reference calling convention: ``x`` (B,T,fs) dequantised floats, ``rnn_state`` a list
        with one (layers,H) tensor or None per slot.  Returns fp32 tensors ``(upsampled, h_n)``.
        LSTM tiers (extension): a slot's state is a ``(h, c)`` pair of (layers,H) tensors and the second return value
        is the pair ``(h_n, c_n)``."""
        b = x.shape[0]
        dev = x.device
        lstm = self.rnn_cell == 'lstm'
        use = torch.tensor([s is not None for s in rnn_state], dtype=torch.uint8).to(dev)

        def dense(pick):
            if not any(s is not None for s in rnn_state):
                return None
            zero = torch.zeros(self.rnn_layers, self.rnn_hidden_size, device=dev)
            return torch.stack([pick(s) if s is not None else zero for s in rnn_state], dim=1).contiguous()

        h_init = self.initial_state(dense((lambda s: s[0]) if lstm else (lambda s: s)), use)
        c_init = self.initial_cell(dense(lambda s: s[1]), use) if lstm else None
        act = torch.float32 if getattr(self, 'precision', 'bf16') == 'fp32' else torch.bfloat16
        upper = upper_conditioning.to(act) if upper_conditioning is not None else None
        up, hn, cn = self._tier(None, 0, None, x.float(), conds.float(), upper, h_init, c_init=c_init)
        return up.float(), ((hn, cn) if lstm else hn)


class SampleLevelLayer(torch.nn.Module):
    """model.py:159-203."""

    def __init__(self, input_samples, conds_size, rnn_hidden_size, q_levels):
        super().__init__()
        self.input_samples = input_samples
        self.q_levels = q_levels
        h, q = rnn_hidden_size, q_levels
        self.emb_layer = _Params(weight=torch.randn(q, q))
        self.emb_layer_expand = _normed(_uniform_(torch.empty(h, q, input_samples),
                                                  math.sqrt(6.0 / (q * input_samples))))     # model.py:177
        bound = 1.0 / math.sqrt(conds_size)
        self.conds_expand = _Params(weight=_uniform_(torch.empty(h, conds_size, 1), bound),
                                    bias=_uniform_(torch.empty(h), bound))
        self.comb_layer = _Params(weight=_uniform_(torch.empty(h, 3 * h), math.sqrt(6.0 / (3 * h))),
                                  bias=torch.zeros(h))                                        # model.py:178-179
        bound = 1.0 / math.sqrt(h)
        self.comb_layer_expand = _normed(_uniform_(torch.empty(h, h, 1), bound), _uniform_(torch.empty(h), bound))
        self.adapt = _normed(_lecun_uniform_(torch.empty(q, h, 1)), torch.zeros(q))          # model.py:180-181

    def _run(self, xs_u8, conds, upper, target_u8, fused):
        fn = SampleLevelFn32 if getattr(self, 'precision', 'bf16') == 'fp32' else SampleLevelFn
        return fn.apply(
            xs_u8, conds, upper, target_u8, fused, self.emb_layer.weight,
            self.emb_layer_expand.weight_g, self.emb_layer_expand.weight_v,
            self.conds_expand.weight, self.conds_expand.bias, self.comb_layer.weight, self.comb_layer.bias,
            self.comb_layer_expand.weight_g, self.comb_layer_expand.weight_v, self.comb_layer_expand.bias,
            self.adapt.weight_g, self.adapt.weight_v, self.adapt.bias)

    def forward(self, x, conds, upper_tier_conditioning):
        """Reference calling convention: ``x`` (B, RF+r0-1) int64 indices -> (B,RF,Q) log-probs."""
        act = torch.float32 if getattr(self, 'precision', 'bf16') == 'fp32' else torch.bfloat16
        return self._run(x.to(torch.uint8).contiguous(), conds.float(), upper_tier_conditioning.to(act), None, False)


class SampleRNNModel(torch.nn.Module):
    """model.py:206-351."""

    def __init__(self, conds_speaker_type, conds_speaker_n, conds_speaker_size, conds_utterance_type,
                 conds_utterance_linguistic_n, conds_utterance_linguistic_emb_size, conds_size, sequence_length, ratios,
                 rnn_layers, rnn_hidden_size, q_type_ulaw, q_levels, fused_loss=False, reference_as_written=False,
                 rnn_cell='gru', precision='bf16'):
        super().__init__()
        if precision not in ('bf16', 'fp32'):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if precision == 'fp32' and rnn_cell != 'gru':
            raise ValueError("precision='fp32' supports GRU tiers only (the reference's cell)")
        self.frame_size = np.prod(ratios)
        self.receptive_field = np.prod(ratios) * sequence_length
        self.quantizer = SampleRNNQuantizer(q_type_ulaw, q_levels)
        self.fused_loss = fused_loss
        self.reference_as_written = reference_as_written
        self.conds_mixer = CondsMixer(conds_speaker_type, conds_speaker_n, conds_speaker_size, conds_utterance_type,
                                      conds_utterance_linguistic_n, conds_utterance_linguistic_emb_size, conds_size)
        self.frames_layers = torch.nn.ModuleList()
        frame_sizes = [int(v) for v in np.cumprod(ratios)]
        for n, fs in enumerate(frame_sizes):
            self.frames_layers.append(FrameLevelLayer(fs, conds_size, ratios[n], rnn_layers[n], rnn_hidden_size[n],
                                                      rnn_cell))
        self.sample_layer = SampleLevelLayer(ratios[0], conds_size, rnn_hidden_size[0], self.quantizer.q_levels)
        self.set_precision(precision)
        self._init_rnn_states(0)

    def set_precision(self, precision):
        """'bf16' (default: bf16 operands, fp32 accumulation) or 'fp32' (fp32-tolerance mode, functional_f32.py)."""
        self.precision = precision
        for mod in [self.conds_mixer, self.sample_layer, *self.frames_layers]:
            mod.precision = precision

    # ---- hidden-state store (model.py:236-250), dense: one (layers,B,H) tensor + validity per tier ----
    def _init_rnn_states(self, batch_size):
        self._state = {n: None for n in range(len(self.frames_layers))}
        self._state_c = {n: None for n in range(len(self.frames_layers))}
        self._state_valid = {n: [False] * batch_size for n in range(len(self.frames_layers))}

    @property
    def rnn_states(self):
        """Reference-shaped view: {layer module: [per-slot (layers,H) tensor or None]}."""
        out = {}
        for n, layer in enumerate(self.frames_layers):
            st, valid = self._state[n], self._state_valid[n]
            out[layer] = [st[:, i, :] if (st is not None and v) else None for i, v in enumerate(valid)]
        return out

    def reset_states(self):
        self._init_rnn_states(0)

    # The carried hidden state is plain module state in the reference (model.py:210,237), so a Skeltorch checkpoint
    # (runner.py:41-45) loses it and a resumed run restarts every stream from ``rnn_h0``.  These two calls let a runner
    # persist it NEXT TO ``state_dict()`` (opt-in: the parameter state_dict keeps exactly the reference's keys).
    def carry_state_dict(self):
        out = {}
        for n in range(len(self.frames_layers)):
            if self._state[n] is not None:
                out[f'frames_layers.{n}.h'] = self._state[n].detach().cpu()
            if self._state_c[n] is not None:
                out[f'frames_layers.{n}.c'] = self._state_c[n].detach().cpu()
            out[f'frames_layers.{n}.valid'] = torch.tensor(self._state_valid[n], dtype=torch.bool)
        return out

    def load_carry_state_dict(self, sd):
        dev = next(self.parameters()).device
        batch = len(sd['frames_layers.0.valid']) if 'frames_layers.0.valid' in sd else 0
        self._init_rnn_states(batch)
        for n in range(len(self.frames_layers)):
            h, c = sd.get(f'frames_layers.{n}.h'), sd.get(f'frames_layers.{n}.c')
            self._state[n] = h.to(dev, torch.float32).contiguous() if h is not None else None
            self._state_c[n] = c.to(dev, torch.float32).contiguous() if c is not None else None
            if f'frames_layers.{n}.valid' in sd:
                self._state_valid[n] = [bool(v) for v in sd[f'frames_layers.{n}.valid'].tolist()]

    def forward(self, x, y, utt_conds, info, reset):
        b, t, _ = utt_conds.size()
        dev = utt_conds.device
        reset_l = [int(r) for r in (reset.tolist() if torch.is_tensor(reset) else reset)]
        if self.reference_as_written or len(self._state_valid[0]) != b:
            self._init_rnn_states(b)                                            # model.py:256-257
        fs_top = self.frames_layers[-1].input_samples
        rf = y.shape[1]

        xq64, xq8 = self.quantizer.quantize_both(x, want_i64=False)            # model.py:260
        yq64, yq8 = self.quantizer.quantize_both(y, want_i64=not self.fused_loss)
        lut = self.quantizer.lut(dev)
        conds = self.conds_mixer(utt_conds, info)                               # model.py:263

        upper = None
        for n in reversed(range(len(self.frames_layers))):                      # model.py:267-276
            layer = self.frames_layers[n]
            valid = self._state_valid[n]
            use = [1 if (r == 0 and valid[i] and self._state[n] is not None) else 0 for i, r in enumerate(reset_l)]
            use_t = torch.tensor(use, dtype=torch.uint8).to(dev, non_blocking=True)
            h_init = layer.initial_state(self._state[n] if any(use) else None, use_t)
            c_init = None
            if layer.rnn_cell == 'lstm':
                c_init = layer.initial_cell(self._state_c[n] if any(use) else None, use_t)
            upper, hn, cn = layer._tier(xq8, fs_top - layer.input_samples, lut, None, conds, upper, h_init,
                                        into_cat=False, c_init=c_init)
            self._state[n] = hn.detach()
            self._state_c[n] = cn.detach() if layer.rnn_cell == 'lstm' else None
            self._state_valid[n] = [r in (0, 1) for r in reset_l]               # model.py:245-250

        r0 = self.sample_layer.input_samples
        xs8 = xq8[:, fs_top - r0:].contiguous()                                 # model.py:279
        keep = [i for i, r in enumerate(reset_l) if r != 2]
        if self.fused_loss:
            logp_t = self.sample_layer._run(xs8, conds, upper, yq8.reshape(-1), True)      # (B, RF)
            if len(keep) != b:
                logp_t = logp_t.index_select(0, torch.tensor(keep, dtype=torch.int64).to(dev))
            return logp_t.unsqueeze(2), torch.zeros(logp_t.shape, dtype=torch.int64, device=dev)
        y_hat = self.sample_layer._run(xs8, conds, upper, yq8.reshape(-1), False)          # (B, RF, Q)
        if len(keep) != b:                                                      # model.py:283-284
            idx = torch.tensor(keep, dtype=torch.int64).to(dev)
            y_hat, yq64 = y_hat.index_select(0, idx), yq64.index_select(0, idx)
        return y_hat, yq64

    def test(self, utt_conds, info, return_logp=False, generator=None, use_graphs=True):
        """model.py:289-351: autoregressive generation.  ``utt_conds`` (B,t,U) - the reference is called
        with B == 1 and ``info`` a single dict; a list of dicts generates B utterances at once.  Returns
        int64 (B, (t+1)*frame_size) whose first frame_size entries are ``quantize_zero()``."""
        from .generate import generate
        self._init_rnn_states(0)                      # model.py:301: generation starts from rnn_h0
        return generate(self, utt_conds, info, return_logp=return_logp, generator=generator, use_graphs=use_graphs)
