"""Counts the tcgen05 / TMA / TMEM / mbarrier SASS instructions per kernel of libdfn.so (cuobjdump -sass; no GPU needed).
python profiles/sass_evidence.py > profiles/sass_evidence_r01.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'dfa-nerf_b200', 'libdfn.so')
KEYS = ('UTCHMMA', 'UTCBAR', 'UBLKCP', 'LDTM', 'STTM', 'SYNCS', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'UTCATOMSWS')

out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
print('SASS evidence (cuobjdump -sass dfa-nerf_b200/libdfn.so, sm_100a): tcgen05 / TMA / TMEM / mbarrier instructions per kernel')
print('(UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy), LDTM/STTM = tcgen05.ld/st, '
      'SYNCS = mbarrier ops, UCGABAR = cluster barrier)\n')
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + '.'):
                counts[fn][k] += 1
for fn, c in counts.items():
    if c:
        print('%-60s %s' % (fn, '  '.join('%s %d' % (k, c[k]) for k in sorted(c))))
