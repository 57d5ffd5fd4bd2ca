"""Blackwell-native evidence from the built library: per kernel of liblsnet_sm100.so the SASS mnemonics that prove
tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA (UTMALDG / UBLKCP) -- B200_PROFILING.md "What proves a Blackwell-native
kernel".  Runs without a GPU.   usage: python tools/sass_counts.py [profiles/r02_sass_counts.txt]"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'lsnet_b200', 'liblsnet_sm100.so')
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, 'profiles', 'r02_sass_counts.txt')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'REDG', 'RED.', 'LDGSTS',
       'FFMA2', 'FHFMA', 'FENCE.VIEW.ASYNC']
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*', '', cur).replace('void ', '')
        counts[cur] = collections.Counter()
        continue
    if cur is None or '/*' not in line:
        continue
    ins = line.split('*/')[1] if '*/' in line else line
    counts[cur]['instructions'] += 1
    for p in pat:
        if (re.search(r'(?<![A-Z])HMMA', ins) if p == 'HMMA' else p in ins):
            counts[cur][p] += 1
with open(dst, 'w') as f:
    f.write('# SASS evidence per kernel of liblsnet_sm100.so (cuobjdump -sass, sm_100a)\n')
    f.write('# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,\n')
    f.write('# SYNCS = mbarrier, FFMA2 = packed fp32x2 FMA, FHFMA = fma.rn.f32.bf16 (fp32 += bf16 x bf16, operands from register halves), REDG = red.global, FENCE.VIEW.ASYNC = fence.proxy.async\n\n')
    for k, c in counts.items():
        tags = ' '.join(f'{p}={c[p]}' for p in pat if c[p])
        f.write(f'{k[:110]:110s} instr={c["instructions"]:6d}  {tags}\n')
        total.update(c)
    f.write('\nTOTAL ' + ' '.join(f'{p}={total[p]}' for p in pat if total[p]) + '\n')
print(open(dst).read()[-2500:])
