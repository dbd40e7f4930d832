"""Per-kernel SASS evidence of the Blackwell-native instructions (B200_PROFILING.md, "What proves a Blackwell-native
kernel"): counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), HMMA
(legacy mma.sync) and LDGSTS (cp.async) in every kernel of libicka_b200.so.  Runs without a GPU.

    python tools/sass_summary.py > profiles/sass_summary_r01.csv
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'icka_b200', 'lib', 'libicka_b200.so')
PATS = OrderedDict([('UTCMMA', r'\bUTC\w*MMA\b'), ('LDTM', r'\bLDTM\b'), ('STTM', r'\bSTTM\b'), ('UTMALDG', r'\bUTMALDG\b'),
                    ('UTMASTG', r'\bUTMASTG\b'), ('UBLKCP', r'\bUBLKCP\b'), ('HMMA', r'\bHMMA\b'), ('LDGSTS', r'\bLDGSTS\b'),
                    ('FFMA', r'\bFFMA\b')])


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True,
                           text=True).stdout.splitlines()
    blocks = re.split(r'Function : \S+', sass)[1:]
    print('# cuobjdump -sass icka_b200/lib/libicka_b200.so (sm_100a), instruction counts per kernel; tools/sass_summary.py')
    print('kernel,instructions,' + ','.join(PATS))
    for name, body in sorted(zip(names, blocks)):
        short = re.sub(r'\(anonymous namespace\)::', '', name)
        short = re.sub(r'\(.*$', '', short)
        c = Counter()
        n = 0
        for line in body.splitlines():
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
            if not m:
                continue
            n += 1
            for k, pat in PATS.items():
                if re.search(pat, m.group(1)):
                    c[k] += 1
        print(f'"{short}",{n},' + ','.join(str(c[k]) for k in PATS))


if __name__ == '__main__':
    sys.exit(main())
