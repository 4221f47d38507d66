"""SASS evidence for profiles/: per kernel of an object file, the Blackwell-specific instruction counts (UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UBLKCP / UTMALDG = TMA bulk / tensor copies, UTCBAR = tcgen05.commit, LDGSTS = cp.async, SYNCS =
mbarrier) and a listing around the first tensor-core instruction.

    python tools/sass_excerpt.py eyoc_b200/csrc/build/sparse_conv_h.o sparse_conv_h_kernel [context]
"""
import re
import subprocess
import sys
from collections import Counter

KEYS = ('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'SYNCS', 'HMMA',
        'FFMA2', 'FADD2', 'FMUL2', 'MUFU', 'POPC', 'SHFL', 'REDUX', 'ATOMS', 'ATOMG', 'RED')


def main(obj, pattern, context=14):
    sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    funcs = re.split(r'\n\s*Function : ', sass)
    for f in funcs[1:]:
        name = f.split('\n', 1)[0].strip()
        if pattern not in name:
            continue
        demangled = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
        ins = re.findall(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)[. ;]', f)
        c = Counter(ins)
        print(f'== {demangled}')
        print(f'   {len(ins)} instructions; ' + ', '.join(f'{k} {c[k]}' for k in KEYS if c[k]))
        lines = [l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4,}\*/\s+\S', l)]
        hit = next((i for i, l in enumerate(lines) if 'UTC' in l and 'MMA' in l), None)
        if hit is None:
            hit = next((i for i, l in enumerate(lines) if any(k in l for k in ('LDTM', 'FFMA2', 'POPC'))), None)
        if hit is not None:
            for l in lines[max(0, hit - context): hit + context]:
                print('   ' + re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l).rstrip())
        print()


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14)
