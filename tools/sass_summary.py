"""SASS mnemonic counts per kernel of libthincurr_b200.so (cuobjdump -sass) + the ptxas -v lines of the tile kernel.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'openfusiontoolkit_b200', 'libthincurr_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout.splitlines()
print('# SASS summary of libthincurr_b200.so (cuobjdump -sass, sm_100a), round 2.  Counts of instruction mnemonics per kernel.')
print('# Evidence: UBLKCP = cp.async.bulk (TMA 1-D) staging, SYNCS = mbarrier, MUFU.RSQ64H = FP64 rsqrt seed, DFMA/DADD/DMUL = FP64 pipe,')
print('# no HMMA/DMMA (tensor cores unused by design: the path is not a contraction), no ATOMG/RED on the matrix in lmat_tile_kernel')
print('# (its ATOMG are the tile-queue counter, the band counters of the streamed build and the optional statistics; ATOMS are shared-memory work counters).')
kern = None
cnt = collections.defaultdict(collections.Counter)
for l in sass:
    m = re.search(r'Function : (\S+)', l)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', l)
    if m and kern:
        op, rest = m.group(1), m.group(2)
        c = cnt[kern]
        c['total'] += 1
        base = op.split('.')[0]
        c[base] += 1
        if op.startswith('MUFU.RSQ64H'):
            c['MUFU.RSQ64H'] += 1
        if base == 'DFMA' and '.reuse' in rest:
            c['DFMA.reuse'] += 1
for k, c in cnt.items():
    if c['total'] < 90:
        continue
    print('\n## %s' % k)
    print('total instructions %d; FP64: DFMA %d (%d with a .reuse operand), DMUL %d, DADD %d; MUFU.RSQ64H %d, MUFU total %d' %
          (c['total'], c['DFMA'], c['DFMA.reuse'], c['DMUL'], c['DADD'], c['MUFU.RSQ64H'], c['MUFU']))
    print(', '.join('%s %d' % (n, c[n]) for n in ('UBLKCP', 'SYNCS', 'BAR', 'LDS', 'STS', 'LDG', 'STG', 'ATOMS', 'ATOMG', 'RED', 'SHFL', 'VOTE', 'HMMA', 'DMMA', 'STL', 'LDL')))
print('\n# ptxas -v (openfusiontoolkit_b200/build.log):')
log = open(os.path.join(ROOT, 'openfusiontoolkit_b200', 'build.log')).read().splitlines()
for i, l in enumerate(log):
    if "Compiling entry function '_ZN3twk16lmat_tile_kernel" in l:
        for j in range(i, min(i + 5, len(log))):
            print('# %d:%s' % (j + 1, log[j]))
