"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    d = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
        d[r[ki][:60]].append(v)
    tot = sum(sum(v) for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:60s} n={len(v):4d} total={sum(v) / 1e3:8.3f} ms  avg={sum(v) / len(v):8.1f} us  {100 * sum(v) / tot:5.1f}%")


if __name__ == '__main__':
    main(sys.argv[1])
