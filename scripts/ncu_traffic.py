"""Extracts per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the captured kernels from
`ncu -i X.ncu-rep --page raw --csv` dumps and writes profiles/r02_ncu_traffic.json, which bench.py reports as
roofline.traffic (captures are taken AT THE BENCHED SIZE: 64 slots, T=4000, H=1024; comb_layer at 1 024 000 rows).

    python scripts/ncu_traffic.py 'gru_kernel<fwd> T=4000'=gpurun_out/gru_fwd_raw.csv comb_layer_fwd=gpurun_out/comb_raw.csv
"""
import csv
import json
import os
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def traffic(path, pick=-1):
    rows = list(csv.reader(ln for ln in open(path, newline='') if ln.startswith('"')))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        tot = 0.0
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(key)
            tot += float(r[i].replace(',', '')) * UNIT[units[i]]
        dur = r[hdr.index('gpu__time_duration.sum')] + ' ' + units[hdr.index('gpu__time_duration.sum')]
        out.append((r[hdr.index('Kernel Name')][:60], tot, dur))
    return out


if __name__ == '__main__':
    res, detail = {}, {}
    for arg in sys.argv[1:]:
        key, path = arg.rsplit('=', 1)
        if not os.path.exists(path):
            continue
        launches = traffic(path)
        if not launches:
            continue
        res[key] = launches[-1][1]                       # the last captured launch (after warm-up)
        detail[key] = [dict(kernel=k, dram_bytes=b, duration=d) for k, b, d in launches]
    res['_detail'] = detail
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r02_ncu_traffic.json')
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))
