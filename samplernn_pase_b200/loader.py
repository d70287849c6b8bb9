"""Sequential-chunk batcher with the reference loader's contract (loader.py:6-84), minus its two
defects: the iterator terminates when every slot is drained (the reference loops forever, loader.py:29-34)
and batches are assembled in pinned host memory so the H2D copy is asynchronous.

Contract kept: ``batch_size`` sticky slots; each slot streams one utterance in chunks of
``x_len = RF + FS - 1`` / ``y_len = RF`` samples and ``L`` conditioning rows, advancing by ``y_len`` /
``L`` per step (loader.py:76-80,83-84); a slot is freed when fewer than L conditioning rows remain
(loader.py:48-50) and refilled - into a random free slot - with ``reset = 1`` (loader.py:55-58); flags are
1 new utterance / 0 continuing / 2 empty slot (zeros, loader.py:67-74); the yielded tuple is
``(x, y, utt_conds, reset, info)`` (loader.py:81).

``dataset`` is any iterable of ``(wav float32 (FS + n*RF,), conds float32 (n*L, U), info dict)`` items that
already carry the FS leading zeros (dataset.py:50) and are truncated to whole chunks (dataset.py:59-62).
"""
import random

import torch


class SequentialChunkLoader:
    def __init__(self, dataset, batch_size, frame_size, sequence_length, conds_width=43, pin_memory=True, seed=None):
        self.dataset = dataset
        self.batch_size = batch_size
        self.frame_size = frame_size
        self.sequence_length = sequence_length
        self.receptive_field = frame_size * sequence_length
        self.conds_width = conds_width
        self.pin_memory = pin_memory and torch.cuda.is_available()
        self.rng = random.Random(seed)

    def iteration_sizes(self):
        """loader.py:83-84."""
        return self.receptive_field + self.frame_size - 1, self.receptive_field, self.sequence_length

    def _alloc(self, *shape):
        t = torch.zeros(*shape)
        return t.pin_memory() if self.pin_memory else t

    def __iter__(self):
        x_len, y_len, l = self.iteration_sizes()
        it = iter(self.dataset)
        slots = [None] * self.batch_size          # [wav, conds, info] per slot
        flags = [None] * self.batch_size          # True new / False continuing / None empty
        exhausted = False
        while True:
            # loader.py:43-50: continuing slots lose the reset flag; drained slots are freed
            for i, item in enumerate(slots):
                if item is None:
                    continue
                flags[i] = False
                if item[1].shape[0] < l:
                    slots[i], flags[i] = None, None
            # loader.py:52-60: refill random free slots
            while not exhausted and any(s is None for s in slots):
                try:
                    wav, conds, info = next(it)
                except StopIteration:
                    exhausted = True
                    break
                if conds.shape[0] < l:
                    continue
                free = [i for i, s in enumerate(slots) if s is None]
                i = self.rng.choice(free)
                slots[i] = [torch.as_tensor(wav, dtype=torch.float32), torch.as_tensor(conds, dtype=torch.float32), info]
                flags[i] = True
            if all(s is None for s in slots):
                return                              # the reference never gets here (loader.py:29-34)
            x = self._alloc(self.batch_size, x_len)
            y = self._alloc(self.batch_size, y_len)
            c = self._alloc(self.batch_size, l, self.conds_width)
            reset = torch.tensor([2 if f is None else int(f) for f in flags])
            info = [s[2] if s is not None else None for s in slots]
            for i, item in enumerate(slots):
                if item is None:
                    continue
                x[i] = item[0][:x_len]                                              # loader.py:76
                y[i] = item[0][self.frame_size:self.frame_size + y_len]             # loader.py:77
                c[i] = item[1][:l]
                item[0] = item[0][y_len:]                                           # loader.py:79-80
                item[1] = item[1][l:]
            yield x, y, c, reset, info
