"""CPU-side checks (``-m "not gpu"``): the C-ABI library loads and exports every symbol the header
declares, the host mirror keeps the reference's module API / state_dict keys, and the product path
fails loudly without CUDA (no fallback)."""
import os
import re

import pytest
import torch

from tests.helpers import Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from samplernn_pase_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'srnn_b200.h')).read()
    declared = set(re.findall(r'\b(srnn_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/srnn_b200.h but not exported'
    assert declared == set(_lib.EXPORTS)


def test_state_dict_keys_and_order_match_reference():
    from samplernn_pase_b200 import SampleRNNModel
    for case, kind in (('gru2_single', 'acoustic'), ('gru2_linguistic', 'linguistic'), ('gru3_multilayer', 'acoustic'),
                       ('gru2_linguistic_lf0', 'linguistic_lf0')):
        g = Golden(case)
        s = g.spec_kwargs()
        m = SampleRNNModel('embedding', int(g.meta['n_spk']), 15, kind, [9, 5, 4, 3], 10, 50, s['sequence_length'],
                           s['ratios'], s['rnn_layers'], s['rnn_hidden_size'], True, 256)
        sd = g.state_dict()
        assert list(m.state_dict().keys()) == list(sd.keys())
        assert all(m.state_dict()[k].shape == v.shape for k, v in sd.items())
        m.load_state_dict(sd)
        assert int(m.frame_size) == int(torch.tensor(s['ratios']).prod())
        assert int(m.receptive_field) == int(m.frame_size) * s['sequence_length']


def test_default_config_parameter_count():
    """SURVEY probe P1: 47 828 534 parameters for config.default.json with 80 speakers."""
    from samplernn_pase_b200 import SampleRNNModel
    m = SampleRNNModel('embedding', 80, 15, 'acoustic', [1, 1, 1, 1], 10, 50, 13, [20, 4], [1, 1], [1024, 1024], True, 256)
    assert sum(p.numel() for p in m.parameters()) == 47_828_534
    assert len(list(m.parameters())) == 44


def test_weight_norm_g_initialised_to_norm_of_v():
    from samplernn_pase_b200 import FrameLevelLayer
    layer = FrameLevelLayer(16, 50, 4, 1, 64)
    v = layer.upsample.weight_v
    assert torch.allclose(layer.upsample.weight_g.view(-1), v.reshape(64, -1).norm(dim=1))
    assert layer.upsample.weight_g.shape == (64, 1, 1)
    assert float(layer.rnn_h0.abs().max()) == 0 and float(layer.upsample_bias.abs().max()) == 0


def test_no_cpu_fallback():
    from samplernn_pase_b200 import SampleRNNQuantizer
    q = SampleRNNQuantizer(True, 256)
    with pytest.raises(RuntimeError, match='CUDA'):
        q.quantize(torch.zeros(4))


def test_dequant_table_is_the_reference_table():
    from samplernn_pase_b200 import SampleRNNQuantizer
    g = Golden('quantizer')
    assert torch.equal(SampleRNNQuantizer(True, 256)._lut_cpu, g.t('dequant_table'))
    assert torch.equal(SampleRNNQuantizer(False, 256)._lut_cpu[:256], g.t('dequant_linear_table'))
    assert SampleRNNQuantizer(True, 256).quantize_zero() == int(g.t('quantize_zero'))


def test_product_does_not_import_the_oracle():
    """No file of the product package imports, loads or executes anything under oracle/."""
    pkg = os.path.join(ROOT, 'samplernn_pase_b200')
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|import_module\([\'"]oracle|[\'"]oracle[/\'"]|#include\s+[<"].*oracle', re.M)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                assert not pat.search(open(os.path.join(root, f)).read()), f


def test_carry_state_dict_round_trip_keeps_reference_keys():
    """The carried hidden state is persisted NEXT TO state_dict() (opt-in): the parameter state_dict keeps exactly the
    reference's keys, and a model restored from (state_dict, carry_state_dict) continues from the same state."""
    from samplernn_pase_b200 import SampleRNNModel
    g = Golden('gru3_multilayer')
    s = g.spec_kwargs()

    def make():
        m = SampleRNNModel('embedding', int(g.meta['n_spk']), 15, 'acoustic', [9, 5, 4, 3], 10, 50, s['sequence_length'],
                           s['ratios'], s['rnn_layers'], s['rnn_hidden_size'], True, 256)
        m.load_state_dict(g.state_dict())
        return m

    a = make()
    keys = list(a.state_dict().keys())
    a._init_rnn_states(3)
    for n, layer in enumerate(a.frames_layers):
        a._state[n] = torch.randn(layer.rnn_layers, 3, layer.rnn_hidden_size)
        a._state_valid[n] = [True, False, True]
    carry = a.carry_state_dict()
    assert list(a.state_dict().keys()) == keys == list(g.state_dict().keys())
    b = make()
    b.load_carry_state_dict(carry)
    for n in range(len(a.frames_layers)):
        assert torch.equal(b._state[n], a._state[n]) and b._state_valid[n] == [True, False, True]
    assert [st is None for st in b.rnn_states[b.frames_layers[0]]] == [False, True, False]


def test_precision_switch_is_validated_and_has_no_cpu_path():
    """``precision`` accepts 'bf16' / 'fp32' only, the fp32-tolerance mode is GRU-only, it keeps the reference's state-dict
    keys, and like the bf16 path it refuses CPU tensors (no fallback)."""
    from samplernn_pase_b200 import SampleRNNModel, ops
    args = ('embedding', 5, 15, 'acoustic', [9, 5, 4, 3], 10, 50, 4, [4, 4], [1, 1], [32, 32], True, 256)
    with pytest.raises(ValueError):
        SampleRNNModel(*args, precision='tf32')
    with pytest.raises(ValueError):
        SampleRNNModel(*args, precision='fp32', rnn_cell='lstm')
    a, b = SampleRNNModel(*args), SampleRNNModel(*args, precision='fp32')
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    assert all(m.precision == 'fp32' for m in [b.conds_mixer, b.sample_layer, *b.frames_layers])
    b.set_precision('bf16')
    assert all(m.precision == 'bf16' for m in [b.conds_mixer, b.sample_layer, *b.frames_layers])
    with pytest.raises(RuntimeError):
        ops.gemm_nt32(torch.zeros(8, 8), torch.zeros(8, 8))
    with pytest.raises(RuntimeError):
        ops.split3(torch.zeros(8, 8), 0)
