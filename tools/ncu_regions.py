"""Per-region summary of an ncu source page: stall samples, executed warp instructions, FP64 / local / shared
instruction counts, grouped by the outermost source line range in tw_lmat.cu (phases of lmat_tile_kernel).

usage: ncu_regions.py <ncu-rep> <lib.so> <kernel name substring>
"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = max([os.path.join(tmp, f) for f in os.listdir(tmp)], key=os.path.getsize)
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l)
off2 = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith('//-----') and '.text.' in l:
        break
    if '//## File' in l:
        # innermost first, then "inlined at" frames outward
        fr = re.findall(r'"([^"]+)", line (\d+)', l)
        cur = [(os.path.basename(f), int(n)) for f, n in fr]
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and cur:
        off2[int(m.group(1), 16)] = (cur, m.group(2))


def func_of(frames):
    """name the phase from the frames (innermost..outermost)"""
    names = []
    for f, n in frames:
        if f == 'tw_device.cuh':
            for lo, hi, nm in DEVREG:
                if lo <= n <= hi:
                    names.append(nm)
                    break
            else:
                names.append('device.cuh:%d' % n)
        elif f == 'tw_lmat.cu':
            for lo, hi, nm in REG:
                if lo <= n <= hi:
                    names.append(nm)
                    break
            else:
                names.append('lmat:%d' % n)
        else:
            names.append(f)
    return names


REG = []
DEVREG = []
_dev = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'openfusiontoolkit_b200', 'csrc', 'tw_device.cuh')).read().splitlines()
_heads = [(i + 1, re.search(r'(\w+)\(', l).group(1)) for i, l in enumerate(_dev) if l.startswith('__device__') and re.search(r'\w+\(', l)]
_heads = [(ln, re.sub(r'^(xmul|xadd|xsub|xdot|xquad)$', 'x-ops', nm)) for ln, nm in _heads]
for k, (ln, nm) in enumerate(_heads):
    DEVREG.append((ln, (_heads[k + 1][0] - 1) if k + 1 < len(_heads) else len(_dev), nm))
src_lines = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'openfusiontoolkit_b200', 'csrc', 'tw_lmat.cu')).read().splitlines()
# regions from "// @region name" markers or function heads
marks = [(i + 1, re.search(r'@region (\S+)', l).group(1)) for i, l in enumerate(src_lines) if '@region' in l]
for k, (ln, nm) in enumerate(marks):
    REG.append((ln, (marks[k + 1][0] - 1) if k + 1 < len(marks) else len(src_lines), nm))

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(src))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[h]
ia, isamp, iex, isrc = H.index('Address'), H.index('# Samples'), H.index('Instructions Executed'), H.index('Source')
stall_cols = [i for i, n in enumerate(H) if n.startswith('stall_') and 'Not Issued' not in n]
base = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= isamp: continue
    addr = int(r[ia], 16)
    if base is None: base = addr
    frames, op = off2.get(addr - base, ([('?', 0)], r[isrc]))
    names = func_of(frames)
    key = names[-1] + (' / ' + names[0] if len(names) > 1 and names[0] != names[-1] else '')
    n, ex = int(r[isamp]), int(r[iex])
    opc = r[isrc].split()[0] if r[isrc] else ''
    if opc.startswith('@'): opc = r[isrc].split()[1]
    c = agg[key]
    c['samples'] += n; c['inst'] += ex
    if opc.startswith(('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX')): c['fp64'] += ex
    if opc.startswith('MUFU'): c['mufu'] += ex
    if opc.startswith(('LDL', 'STL')): c['local'] += ex
    if opc.startswith(('LDS', 'STS', 'ATOMS')): c['shared'] += ex
    if opc.startswith(('LDG', 'STG', 'LD.', 'ST.')): c['global'] += ex
    if opc.startswith('BAR'): c['bar'] += ex
    for cc in stall_cols:
        v = int(r[cc])
        if v: c[H[cc]] += v
    tot['samples'] += n; tot['inst'] += ex
print('total samples %d, warp instructions %d' % (tot['samples'], tot['inst']))
print('%-44s %7s %7s %7s %6s %6s %6s %6s  top stalls' % ('region', 'samp%', 'inst%', 'fp64%', 'mufu%', 'lds%', 'loc%', 'glob%'))
for k, c in sorted(agg.items(), key=lambda kv: -kv[1]['samples']):
    if c['samples'] < 0.002 * tot['samples']: continue
    st = ', '.join('%s %.0f%%' % (a.replace('stall_', ''), 100.0 * b / c['samples']) for a, b in collections.Counter({a: b for a, b in c.items() if a.startswith('stall_')}).most_common(4))
    print('%-44s %6.2f%% %6.2f%% %6.1f%% %5.1f%% %5.1f%% %5.1f%% %5.1f%%  %s' % (k, 100.0 * c['samples'] / tot['samples'], 100.0 * c['inst'] / tot['inst'],
          100.0 * c['fp64'] / max(c['inst'], 1), 100.0 * c['mufu'] / max(c['inst'], 1), 100.0 * c['shared'] / max(c['inst'], 1), 100.0 * c['local'] / max(c['inst'], 1),
          100.0 * c['global'] / max(c['inst'], 1), st))
