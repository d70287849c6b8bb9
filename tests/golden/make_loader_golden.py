"""Generates tests/golden/loader_schedule.npz from the REFERENCE loader (run in the build container only:
imports /root/reference).  A fake dataset with the attributes `SampleRNNPASELoader` reads (loader.py:16-27,37-38:
frame_size, sequence_length, conds_utterance_type, shuffle_utterances, __getitem__/__len__) feeds it; the global
`random` generator is seeded, two epochs are recorded (the reference iterator never ends, loader.py:29-34, so an
epoch is cut when every slot is empty)."""
import os
import random
import sys

import numpy as np

sys.path.insert(0, '/root/reference')
from samplernn_pase.loader import SampleRNNPASELoader  # noqa: E402

FS, L, WIDTH, BATCH, SEED = 4, 3, 43, 3, 7
LENGTHS = [2, 1, 3, 1, 4, 2, 1]            # chunks per utterance


def make_item(u, n):
    wav = np.concatenate([np.zeros(FS, dtype=np.float32), (np.arange(n * FS * L, dtype=np.float32) + 1000 * u) / 8192])
    conds = np.full((n * L, WIDTH), float(u), dtype=np.float32) + np.arange(n * L, dtype=np.float32)[:, None] / 64
    return wav, conds, {'speaker': {'index': u}, 'utt': u}


class FakeDataset:
    frame_size = FS
    sequence_length = L
    conds_utterance_type = 'acoustic'

    def __init__(self):
        self.utterances_ids = list(range(len(LENGTHS)))

    def __getitem__(self, item):
        u = self.utterances_ids[item]                      # dataset.py:39-40 (IndexError ends the iteration)
        return make_item(u, LENGTHS[u])

    def __len__(self):
        return len(self.utterances_ids)

    def shuffle_utterances(self):
        random.shuffle(self.utterances_ids)                # dataset.py:56-57


def main():
    random.seed(SEED)
    loader = SampleRNNPASELoader(FakeDataset(), BATCH)
    out = dict(fs=FS, l=L, width=WIDTH, batch=BATCH, seed=SEED, lengths=np.array(LENGTHS))
    for epoch in range(2):
        k = 0
        for x, y, c, reset, info in loader:
            if bool((reset == 2).all()):
                break
            pre = f'e{epoch}.s{k}.'
            out[pre + 'x'], out[pre + 'y'], out[pre + 'c'] = x.numpy(), y.numpy(), c.numpy()
            out[pre + 'reset'] = reset.numpy()
            out[pre + 'utt'] = np.array([-1 if i is None else i['utt'] for i in info])
            k += 1
        out[f'e{epoch}.steps'] = k
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'loader_schedule.npz')
    np.savez_compressed(path, **out)
    print(path, {k: out[k] for k in ('e0.steps', 'e1.steps')})


if __name__ == '__main__':
    main()
