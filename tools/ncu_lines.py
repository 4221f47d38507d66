"""Where a profiled kernel spends its instructions: `ncu -i REP --page source --csv` (SASS view) -> opcode histogram of the
executed warp instructions and the hottest instructions in address order (needs a capture with --set full)."""
import collections
import csv
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', *sys.argv[3:]], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next((i for i, r in enumerate(rows) if any(c.strip() == 'Source' for c in r) and len(r) > 4), None)
    if hdr is None:
        print(out[:2000])
        return
    h = [c.strip() for c in rows[hdr]]
    print('columns:', h)
    ci = next((i for i, c in enumerate(h) if c == '# Warp Instructions Executed'), None)
    if ci is None:
        ci = next(i for i, c in enumerate(h) if 'Instructions Executed' in c)
    si = h.index('Source')
    ai = next((i for i, c in enumerate(h) if c in ('Address', '#')), 0)
    data = []
    for r in rows[hdr + 1:]:
        try:
            data.append((float(r[ci].replace(',', '') or 0), r[ai], r[si].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1.0
    print(f'total executed warp instructions {tot:.4e} over {len(data)} SASS instructions')
    ops = collections.Counter()
    for n, _, src in data:
        toks = src.replace('@', ' @').split()
        op = next((t for t in toks if not t.startswith('@') and not t.startswith('!')), '?')
        ops[op.split('.')[0]] += n
    print('--- opcode histogram')
    for op, n in ops.most_common(25):
        print(f'{100 * n / tot:5.1f}%  {op}')
    print('--- cumulative share by position (deciles of the listing)')
    acc, step = 0.0, max(1, len(data) // 20)
    for i in range(0, len(data), step):
        part = sum(d[0] for d in data[i:i + step])
        acc += part
        print(f'  instr {i:5d}..{min(len(data), i + step) - 1:5d}  {100 * part / tot:5.1f}%   first: {data[i][2][:90]}')
    if '# Samples' in h:
        smi = h.index('# Samples')
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
        samp = []
        for r in rows[hdr + 1:]:
            try:
                n = float(r[smi].replace(',', '') or 0)
            except (ValueError, IndexError):
                continue
            st = sorted(((float(r[i].replace(',', '') or 0), c) for i, c in stall_cols if i < len(r)), reverse=True)[:2]
            samp.append((n, r[ai], r[si].strip(), st))
        ts = sum(x[0] for x in samp) or 1.0
        agg = collections.Counter()
        for r in rows[hdr + 1:]:
            for i, c in stall_cols:
                try:
                    agg[c] += float(r[i].replace(',', '') or 0)
                except (ValueError, IndexError):
                    pass
        print('--- stall reasons (all samples):', ', '.join(f'{c} {100 * n / max(1.0, sum(agg.values())):.1f}%' for c, n in agg.most_common(8)))
        print(f'--- top {top} instructions by stall samples ({ts:.0f} samples)')
        for n, addr, src, st in sorted(samp, reverse=True)[:top]:
            print(f'{100 * n / ts:5.2f}%  {addr[-6:]:>6s}  {src[:90]:90s} {st[0][1] if st else ""} {st[1][1] if len(st) > 1 else ""}')
    print(f'--- top {top} instructions')
    for n, addr, src in sorted(data, reverse=True)[:top]:
        print(f'{100 * n / tot:5.2f}%  {addr:>8s}  {src[:120]}')


if __name__ == '__main__':
    main()
