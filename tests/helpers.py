"""Shared helpers for the parity tests: golden-file access and error metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Golden:
    """One ``tests/golden/*.npz`` file produced by ``make_golden.py`` from the reference."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'))
        self.meta = {k[5:]: self.z[k] for k in self.z.files if k.startswith('meta.')}

    def state_dict(self):
        return {k[3:]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith('sd.')}

    def t(self, key):
        return torch.from_numpy(np.asarray(self.z[key]))

    def has(self, key):
        return key in self.z.files

    def chunk(self, k):
        pre = f'c{k}.'
        return {key[len(pre):]: torch.from_numpy(np.asarray(self.z[key])) for key in self.z.files if key.startswith(pre)}

    @property
    def chunks(self):
        return int(self.meta['chunks'])

    def spec_kwargs(self):
        m = self.meta
        return dict(ratios=[int(v) for v in m['ratios']], rnn_layers=[int(v) for v in m['layers']],
                    rnn_hidden_size=[int(m['hidden'])] * len(m['ratios']), sequence_length=int(m['seq_len']),
                    conds_utterance_type=str(m['kind']))


def rel_l2(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
