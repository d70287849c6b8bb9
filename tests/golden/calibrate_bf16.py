"""Calibrates the bf16 tolerance of tests/test_gpu_model.py: runs the UNMODIFIED reference on the
golden inputs under torch's own CPU bf16 autocast and reports its error against its fp32 self.
(The reference has no mixed-precision mode; autocast is the closest "reference in bf16".)

    python tests/golden/calibrate_bf16.py
"""
import sys
import warnings

import torch

sys.path.insert(0, '/root/reference')
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
warnings.filterwarnings('ignore')
import make_golden as mg                      # noqa: E402
from tests.helpers import Golden, cosine, rel_l2   # noqa: E402


def run(name):
    g = Golden(name)
    kw = g.spec_kwargs()
    m = mg.build(kw['ratios'], kw['rnn_layers'], kw['rnn_hidden_size'][0], kw['sequence_length'],
                 kw['conds_utterance_type'], int(g.meta['n_spk']))
    m.load_state_dict(g.state_dict())
    c = g.chunk(0)
    info = [None if int(r) == 2 else {'speaker': {'index': int(s)}} for s, r in zip(c['speakers'], c['reset'])]
    m.zero_grad()
    with torch.autocast('cpu', dtype=torch.bfloat16):
        y_hat, yq = m(c['x'], c['y'], c['conds'], info, c['reset'])
    loss = torch.nn.functional.nll_loss(y_hat.float().view(-1, 256), yq.view(-1))
    loss.backward()
    worst_r, worst_c, who = 0.0, 1.0, ''
    for pn, p in m.named_parameters():
        ref = c['grad.' + pn]
        if float(ref.norm()) == 0 or p.grad is None:
            continue
        r, cs = rel_l2(p.grad, ref), cosine(p.grad, ref)
        if r > worst_r:
            worst_r, worst_c, who = r, cs, pn
    dl = float((y_hat.float() - c['y_hat']).abs().max())
    print(f'{name}: loss rel {abs(float(loss) - float(c["loss"])) / float(c["loss"]):.2e}  max|dlogp| {dl:.2e}  '
          f'worst grad rel_l2 {worst_r:.3e} cos {worst_c:.5f} ({who})')


if __name__ == '__main__':
    for n in ['gru2_single', 'gru2_carry', 'gru2_aswritten', 'gru3_multilayer', 'gru2_linguistic', 'gru2_default_ratios']:
        run(n)
