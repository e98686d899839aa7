"""Join ncu's per-SASS-instruction counts (--page source --csv) with nvdisasm --print-line-info to get
executed warp-instructions and stall samples per CUDA source line (development aid)."""
import csv, re, sys, collections
dis, srccsv, kernel_pat = sys.argv[1], sys.argv[2], sys.argv[3]
# parse disassembly of the wanted kernel
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and re.search(kernel_pat, l))
addr2line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//----'):
        if addr2line: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', l)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; data = rows[2:]
ia, iex, ismp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in data:
    try: a = int(r[ia], 16); n = int(r[iex]); s = int(r[ismp])
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base, ('?', 0))
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += 1; tot += n
tots = sum(v[1] for v in agg.values())
print('total warp-instr', tot, 'samples', tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print('%-28s line %4d  instr %9d %5.1f%%  samples %6d %5.1f%%  static %4d' % (k[0], k[1], v[0], 100 * v[0] / tot, v[1], 100 * v[1] / max(tots, 1), v[2]))
