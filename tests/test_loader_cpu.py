"""The sequential-chunk loader keeps the reference loader's schedule (loader.py:29-84) and terminates."""
import torch

from samplernn_pase_b200.loader import SequentialChunkLoader


def make_dataset(lengths, fs, l, width=43):
    items = []
    for u, n in enumerate(lengths):
        wav = torch.cat([torch.zeros(fs), torch.arange(n * fs * l, dtype=torch.float32) + 1000 * u])
        conds = torch.full((n * l, width), float(u))
        items.append((wav, conds, {'speaker': {'index': u}, 'utt': u}))
    return items


def test_schedule_flags_and_termination():
    fs, l = 4, 3
    rf = fs * l
    ds = make_dataset([2, 1, 3], fs, l)
    loader = SequentialChunkLoader(ds, batch_size=2, frame_size=fs, sequence_length=l, pin_memory=False, seed=0)
    seen = {u: [] for u in range(3)}
    steps = list(loader)
    assert 3 <= len(steps) <= 6                      # finite: the reference iterator never ends
    slot_owner = [None, None]
    for x, y, c, reset, info in steps:
        assert x.shape == (2, rf + fs - 1) and y.shape == (2, rf) and c.shape == (2, l, 43)
        for i in range(2):
            r = int(reset[i])
            if r == 2:
                assert info[i] is None and float(x[i].abs().sum()) == 0 and float(c[i].abs().sum()) == 0
                slot_owner[i] = None
                continue
            u = info[i]['utt']
            if r == 1:
                assert slot_owner[i] != u           # a new utterance entered this slot
                slot_owner[i] = u
            else:
                assert slot_owner[i] == u           # sticky slot: the hidden state of a slot stays meaningful
            seen[u].append((x[i].clone(), y[i].clone()))
            assert torch.equal(x[i, fs:], y[i, :rf - 1])     # x/y overlap (loader.py:76-77)
    for u, n in enumerate([2, 1, 3]):
        assert len(seen[u]) == n                     # every chunk of every utterance exactly once, in order
        wav = ds[u][0]
        for k, (x, y) in enumerate(seen[u]):
            assert torch.equal(x, wav[k * rf: k * rf + rf + fs - 1])
            assert torch.equal(y, wav[fs + k * rf: fs + (k + 1) * rf])


def test_first_chunk_has_reset_one_and_leading_zeros():
    fs, l = 4, 2
    ds = make_dataset([2], fs, l)
    x, y, c, reset, info = next(iter(SequentialChunkLoader(ds, 1, fs, l, pin_memory=False)))
    assert reset.tolist() == [1] and float(x[0, :fs].abs().sum()) == 0


def test_replays_the_reference_loader_schedule_exactly():
    """tests/golden/loader_schedule.npz was recorded from the imported reference loader (make_loader_golden.py:
    loader.py:29-84 + dataset.shuffle_utterances, global ``random.seed(7)``, two epochs).  With ``shuffle=True`` and
    the same seed this loader must yield the same batches, reset flags and slot occupancy, step by step, and - unlike
    the reference - stop at the end of each epoch."""
    import os

    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'loader_schedule.npz'))
    fs, l, width, batch, seed = (int(z[k]) for k in ('fs', 'l', 'width', 'batch', 'seed'))
    lengths = [int(v) for v in z['lengths']]

    class Dataset:
        def __len__(self):
            return len(lengths)

        def __getitem__(self, u):
            n = lengths[u]
            wav = np.concatenate([np.zeros(fs, dtype=np.float32),
                                  (np.arange(n * fs * l, dtype=np.float32) + 1000 * u) / 8192])
            conds = np.full((n * l, width), float(u), dtype=np.float32) + np.arange(n * l, dtype=np.float32)[:, None] / 64
            return wav, conds, {'speaker': {'index': u}, 'utt': u}

    loader = SequentialChunkLoader(Dataset(), batch, fs, l, conds_width=width, pin_memory=False, seed=seed, shuffle=True)
    for epoch in range(2):
        steps = list(loader)
        assert len(steps) == int(z[f'e{epoch}.steps'])
        for k, (x, y, c, reset, info) in enumerate(steps):
            pre = f'e{epoch}.s{k}.'
            assert reset.tolist() == z[pre + 'reset'].tolist(), (epoch, k)
            assert [-1 if i is None else i['utt'] for i in info] == z[pre + 'utt'].tolist(), (epoch, k)
            assert np.array_equal(x.numpy(), z[pre + 'x']) and np.array_equal(y.numpy(), z[pre + 'y'])
            assert np.array_equal(c.numpy(), z[pre + 'c'])
