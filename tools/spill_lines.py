"""List where local-memory (spill) instructions of a kernel come from: source line -> STL/LDL count.
usage: spill_lines.py <lib.so> <kernel name substring>"""
import os, re, subprocess, sys, tempfile, collections
lib, kern = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = max([os.path.join(tmp, f) for f in os.listdir(tmp)], key=os.path.getsize)
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l)
cur = None
cnt = collections.Counter()
for l in dis[start + 1:]:
    if l.startswith('//-----') and '.text.' in l:
        break
    if '//## File' in l:
        fr = re.findall(r'"([^"]+)", line (\d+)', l)
        cur = ' <- '.join('%s:%s' % (os.path.basename(f), n) for f, n in fr[:3])
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
    if m and re.search(r'\b(STL|LDL)\b', m.group(1)):
        cnt[(cur, 'STL' if 'STL' in m.group(1) else 'LDL')] += 1
for (k, op), n in cnt.most_common(40):
    print('%4d %s  %s' % (n, op, k))
