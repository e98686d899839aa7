"""Executed warp-instructions and stall samples per SOURCE LINE (ncu --page source --csv joined with
nvdisasm --print-line-info); development aid.
usage: ncu_lines.py dis.txt src.csv <kernel regex> [top]"""
import collections
import csv
import re
import sys

dis, srccsv, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and re.search(pat, l))
addr2line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith('.text.') and addr2line:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', l)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
data = rows[2:]
ia, iex, ismp, isrc = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
ops = collections.defaultdict(int)
for r in data:
    try:
        a = int(r[ia], 16); n = int(r[iex]); s = int(r[ismp])
    except Exception:
        continue
    if base is None:
        base = a
    k = addr2line.get(a - base, ('?', 0))
    agg[k][0] += n; agg[k][1] += s; agg[k][2] += 1
    op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
    ops[op.split('.')[0]] += n
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total executed', tot, 'samples', ts, 'static', len(data))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-28s instr %5.1f%%  samples %5.1f%%  static %4d' % ('%s:%d' % k, 100 * v[0] / tot, 100 * v[1] / ts, v[2]))
print('--- opcodes')
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
    print('%-10s %5.1f%%' % (k, 100 * v / tot))
