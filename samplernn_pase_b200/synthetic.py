"""Synthetic workload with the reference loader's chunk schedule (SURVEY.md 8(d) "Synthetic inputs";
loader.py:76-84, dataset.py:50): used by bench.py and the profiling scripts so that the product path never
touches ``oracle/``."""
import torch


def synthetic_utterances(frame_size, receptive_field, seq_len, batch, chunks, seed=4321, conds_width=43,
                         n_speakers=126):
    """wav ~ U(-0.99, 0.99) with ``frame_size`` leading zeros per utterance, conds ~ N(0,1), speaker i % n."""
    g = torch.Generator().manual_seed(seed)
    wav = (torch.rand(batch, frame_size + chunks * receptive_field, generator=g) * 2 - 1) * 0.99
    wav[:, :frame_size] = 0.0
    conds = torch.randn(batch, chunks * seq_len, conds_width, generator=g)
    speakers = torch.arange(batch) % n_speakers
    return wav, conds, speakers


def chunk_of(frame_size, receptive_field, seq_len, wav, conds, k):
    """x = wav[k*RF : k*RF + RF + FS - 1], y = wav[FS + k*RF : FS + (k+1)*RF], conds rows [k*L, (k+1)*L)."""
    fs, rf, l = frame_size, receptive_field, seq_len
    x = wav[:, k * rf: k * rf + rf + fs - 1]
    y = wav[:, fs + k * rf: fs + (k + 1) * rf]
    return x.contiguous(), y.contiguous(), conds[:, k * l:(k + 1) * l].contiguous()
