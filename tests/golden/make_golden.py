"""Generate golden vectors by running the UNMODIFIED reference (imported from
/root/reference) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, the ``.npz`` files it produces can.  Each file
holds the reference ``state_dict`` (same key names), the inputs and the reference outputs
(log-probs, targets, loss, every parameter gradient, carried hidden states).

The only interventions on the reference are the ones SURVEY.md 8(c) documents:
  * O-B "carry": ``model._init_rnn_states`` is called once and then neutralised, which
    removes the effect of the ``hasattr(self, 'rnnstates')`` typo (model.py:256);
  * zero-initialised tensors are perturbed (they would hide bugs).
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')
import samplernn_pase.model as ref_model      # noqa: E402
import samplernn_pase.utils as ref_utils      # noqa: E402
import samplernn_pase.optimizer as ref_opt    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build(ratios, layers, hidden, seq_len, kind='acoustic', n_spk=7, ling_n=(9, 5, 4, 3), seed=1234):
    torch.manual_seed(seed)
    m = ref_model.SampleRNNModel(
        conds_speaker_type='embedding', conds_speaker_n=n_spk, conds_speaker_size=15,
        conds_utterance_type=kind, conds_utterance_linguistic_n=list(ling_n),
        conds_utterance_linguistic_emb_size=10, conds_size=50, sequence_length=seq_len, ratios=list(ratios),
        rnn_layers=list(layers), rnn_hidden_size=[hidden] * len(ratios), q_type_ulaw=True, q_levels=256)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, prm in m.named_parameters():
            if float(prm.abs().max()) == 0.0:
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
    return m


def make_inputs(m, batch, chunks, kind, seed=4321, ling_n=(9, 5, 4, 3)):
    g = torch.Generator().manual_seed(seed)
    fs, rf = int(m.frame_size), int(m.receptive_field)
    l = rf // fs
    width = {'acoustic': 43, 'linguistic': 55, 'linguistic_lf0': 57}[kind]
    wav = (torch.rand(batch, fs + chunks * rf, generator=g) * 2 - 1) * 0.99
    wav[:, :fs] = 0.0
    conds = torch.randn(batch, chunks * l, width, generator=g)
    if kind != 'acoustic':
        cats = ([2, 3, 4, 5, 6], [27], [31, 33, 41], [49])
        for cols, ncat in zip(cats, ling_n):
            for c in cols:
                conds[:, :, c] = torch.randint(0, ncat, (batch, chunks * l), generator=g).float()
    return wav, conds


def chunk(m, wav, conds, k):
    fs, rf = int(m.frame_size), int(m.receptive_field)
    l = rf // fs
    return (wav[:, k * rf: k * rf + rf + fs - 1].contiguous(), wav[:, fs + k * rf: fs + (k + 1) * rf].contiguous(),
            conds[:, k * l:(k + 1) * l].contiguous())


def info_for(batch, reset, n_spk):
    return [None if r == 2 else {'speaker': {'index': (3 * i + 1) % n_spk}} for i, r in enumerate(reset)]


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(name, os.path.getsize(path) // 1024, 'KiB')


def quantizer_case():
    q = ref_utils.SampleRNNQuantizer(True, 256)
    g = torch.Generator().manual_seed(7)
    x = torch.cat([torch.linspace(-0.99, 0.99, 20001), (torch.rand(20000, generator=g) * 2 - 1) * 0.99,
                   torch.tensor([0.0, -0.0, 1e-8, -1e-8, 0.5, -0.5, 0.9899999, -0.9899999])])
    idx = q.quantize(x)
    table = q.dequantize(torch.arange(257))
    ql = ref_utils.SampleRNNQuantizer(False, 256)
    xl = (torch.rand(4001, generator=g) * 2 - 1) * 0.7
    save('quantizer', x=x, idx=idx, dequant_table=table, x_linear=xl, idx_linear=ql.quantize(xl),
         dequant_linear_table=ql.dequantize(torch.arange(256)), quantize_zero=np.int64(q.quantize_zero()))


def model_case(name, ratios, layers, hidden, seq_len, batch, resets, kind='acoustic', carry=True, n_spk=7):
    """``resets``: list over chunks of per-slot reset flags.  Saves per-chunk outputs and the
    gradients of the summed per-chunk losses' LAST chunk (one backward per chunk, like training)."""
    m = build(ratios, layers, hidden, seq_len, kind, n_spk)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    wav, conds = make_inputs(m, batch, len(resets), kind)
    if carry:                                   # O-B: neutralise the typo'd re-initialisation
        m._init_rnn_states(batch)
        m._init_rnn_states = lambda b: None
    out = {}
    for k, reset in enumerate(resets):
        x, y, c = chunk(m, wav, conds, k)
        for i, r in enumerate(reset):           # loader.py:70-74: empty slots are all-zero
            if r == 2:
                x[i].zero_(); y[i].zero_(); c[i].zero_()
        info = info_for(batch, reset, n_spk)
        m.zero_grad()
        y_hat, yq = m(x, y, c, info, torch.tensor(reset))
        loss = torch.nn.functional.nll_loss(y_hat.view(-1, y_hat.size(2)), yq.view(-1))
        loss.backward()
        out[f'c{k}.x'] = x; out[f'c{k}.y'] = y; out[f'c{k}.conds'] = c
        out[f'c{k}.reset'] = np.asarray(reset, dtype=np.int64)
        out[f'c{k}.speakers'] = np.asarray([0 if it is None else it['speaker']['index'] for it in info], dtype=np.int64)
        out[f'c{k}.y_hat'] = y_hat; out[f'c{k}.yq'] = yq; out[f'c{k}.loss'] = loss
        if k in (0, len(resets) - 1):           # gradients of the first and last chunk only (file size)
            for pname, prm in m.named_parameters():
                # a parameter that received no gradient (rnn_h0 when every slot carries) is stored as zeros
                out[f'c{k}.grad.{pname}'] = torch.zeros_like(prm) if prm.grad is None else prm.grad.clone()
        for n, layer in enumerate(m.frames_layers):
            st = m.rnn_states[layer]
            h = torch.stack([s if s is not None else torch.full((layers[n], hidden), float('nan')) for s in st], dim=1)
            out[f'c{k}.state.{n}'] = h
    meta = dict(ratios=np.asarray(ratios), layers=np.asarray(layers), hidden=np.int64(hidden),
                seq_len=np.int64(seq_len), chunks=np.int64(len(resets)), carry=np.int64(int(carry)),
                kind=np.asarray(kind), n_spk=np.int64(n_spk))
    save(name, **{'sd.' + k: v for k, v in sd.items()}, **out, **{'meta.' + k: v for k, v in meta.items()})


def adam_case():
    torch.manual_seed(3)
    w = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(11))]
    w0 = [p.detach().clone() for p in w]
    opt = ref_opt.AdamClipped(w, lr=1e-3)
    grads = []
    for s in range(3):
        gs = [3.0 * torch.randn_like(p) for p in w]       # scaled so the [-1,1] clamp is active
        grads.append(gs)
        for p, g in zip(w, gs):
            p.grad = g.clone()
        opt.step()
    save('adam_clipped', w0_a=w0[0], w0_b=w0[1], w3_a=w[0], w3_b=w[1],
         **{f'g{s}_{n}': grads[s][i] for s in range(3) for i, n in enumerate('ab')})


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'lf0':          # add one case without rewriting the committed fixtures
        model_case('gru2_linguistic_lf0', [4, 2], [1, 1], 32, 2, 2, [[1, 1], [0, 1]], kind='linguistic_lf0')
        sys.exit(0)
    quantizer_case()
    adam_case()
    model_case('gru2_single', [4, 2], [1, 1], 32, 3, 3, [[1, 1, 1]])
    model_case('gru2_carry', [4, 2], [1, 1], 32, 3, 4, [[1, 1, 1, 2], [0, 0, 1, 2], [0, 1, 0, 1], [2, 0, 0, 0]])
    model_case('gru2_aswritten', [4, 2], [1, 1], 32, 3, 3, [[1, 1, 1], [0, 0, 0]], carry=False)
    model_case('gru3_multilayer', [5, 2, 3], [1, 2, 1], 32, 2, 2, [[1, 1], [0, 0]])
    model_case('gru2_linguistic', [4, 2], [1, 1], 32, 2, 2, [[1, 1]], kind='linguistic')
    model_case('gru2_default_ratios', [20, 4], [1, 1], 16, 2, 2, [[1, 1], [0, 0]])
    model_case('gru2_linguistic_lf0', [4, 2], [1, 1], 32, 2, 2, [[1, 1], [0, 1]], kind='linguistic_lf0')
