"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv [--last-step-launches N]
"""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = r['Kernel Name']
        name = re.sub(r'\(.*$', '', name)
        name = re.sub(r'^void ', '', name)
        name = name.replace('(anonymous namespace)::', '')
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
        rows.append((name, ns))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for n, ns in rows:
        tot[n] += ns
        cnt[n] += 1
    total = sum(tot.values())
    lines = [f'# {path}: {len(rows)} launches, {total / 1e6:.3f} ms total (cold-cache, serialised: compare shares)',
             f'{"kernel":90s} {"launches":>8s} {"total_ms":>10s} {"share":>7s} {"avg_us":>10s}']
    for n in sorted(tot, key=tot.get, reverse=True)[:40]:
        lines.append(f'{n[:90]:90s} {cnt[n]:8d} {tot[n] / 1e6:10.3f} {100 * tot[n] / total:6.1f}% {tot[n] / cnt[n] / 1e3:10.1f}')
    text = '\n'.join(lines)
    print(text)
    if out:
        open(out, 'w').write(text + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
