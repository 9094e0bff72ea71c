"""Join an ncu SASS source page (CSV) with nvdisasm -g line info: stall samples per CUDA source line.

usage: ncu_lines.py <ncu-rep> <lib.so> <kernel mangled-name substring> [top]
"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = max([os.path.join(tmp, f) for f in os.listdir(tmp)], key=os.path.getsize)
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
# locate kernel section
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l)
off2line = {}
cur = None
stack = ''
for l in dis[start + 1:]:
    if l.startswith('//-----') and '.text.' in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        stack = ' <- %s:%s' % (os.path.basename(inl.group(1)), inl.group(2)) if inl else ''
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and cur:
        off2line[int(m.group(1), 16)] = (cur, stack, m.group(2))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(src))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[h]
ia, isamp, iex = H.index('Address'), H.index('# Samples'), H.index('Instructions Executed')
stall_cols = [i for i, n in enumerate(H) if n.startswith('stall_') and 'Not Issued' not in n]
base = None
per_line = collections.Counter(); per_line_inst = collections.Counter(); per_line_stall = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[h + 1:]:
    if len(r) <= isamp: continue
    addr = int(r[ia], 16)
    if base is None: base = addr
    off = addr - base
    n = int(r[isamp]); tot += n
    key = off2line.get(off, (('?', 0), '', ''))
    k = '%s:%d%s' % (key[0][0], key[0][1], key[1])
    per_line[k] += n
    per_line_inst[k] += int(r[iex])
    for c in stall_cols:
        v = int(r[c])
        if v: per_line_stall[k][H[c]] += v
print('total samples', tot)
for k, n in per_line.most_common(top):
    st = ', '.join('%s %d' % (a.replace('stall_', ''), b) for a, b in per_line_stall[k].most_common(3))
    print('%6.2f%%  inst %12d  %-60s %s' % (100.0 * n / tot, per_line_inst[k], k, st))
