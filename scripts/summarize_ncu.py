"""Turns ncu CSV output into the short text summaries kept under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv   # `ncu --metrics gpu__time_duration.sum --csv` launch list
  python scripts/summarize_ncu.py full gpurun_out/hot_raw.csv        # `ncu -i x.ncu-rep --page raw --csv` of a --set full capture
"""
import collections
import csv
import re
import sys

KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'launch__shared_mem_per_block_dynamic']


def rows_of(path):
    lines = [ln for ln in open(path, newline='') if ln.startswith('"')]
    return list(csv.reader(lines))


def short(name):
    name = re.sub(r'\(.*', '', name)
    return name.replace('srnn::', '').replace('at::native::', 'at::')[:110]


def launches(path):
    rows = rows_of(path)
    hdr = rows[0]
    k, v, u = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= v or not r[v]:
            continue
        ns = float(r[v].replace(',', '')) * {'ns': 1.0, 'us': 1e3, 'ms': 1e6}.get(r[u], 1.0)
        d = agg.setdefault(short(r[k]), [0, 0.0])
        d[0] += 1
        d[1] += ns
    total = sum(d[1] for d in agg.values())
    print(f'launches {sum(d[0] for d in agg.values())}  total {total / 1e6:.1f} ms')
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f'{ns / 1e6:10.3f} ms {100 * ns / total:5.1f}% x{n:<5d} {name}')


def full(path):
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    k = hdr.index('Kernel Name')
    cols = [(key, next((i for i, h in enumerate(hdr) if h == key or h.endswith('.' + key)), None)) for key in KEYS]
    for n, r in enumerate(rows[2:], 1):
        print(f'[{n}] {short(r[k])}')
        for key, i in cols:
            if i is not None and i < len(r) and r[i] != '':
                print(f'    {key} = {r[i]} {units[i]}')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
