"""Executed warp-instructions and stall samples per source function (ncu --page source --csv joined with
nvdisasm --print-line-info); development aid."""
import csv, re, collections, sys
dis, srccsv, pat, nwarps, nrounds = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and re.search(pat, l))
addr2line = {}; cur = None
for l in lines[start + 1:]:
    if l.startswith('.text.') and addr2line: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', l)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv))); hdr = rows[1]; data = rows[2:]
ia, iex, ismp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = None; agg = collections.defaultdict(lambda: [0, 0])
for r in data:
    try: a = int(r[ia], 16); n = int(r[iex]); s = int(r[ismp])
    except Exception: continue
    if base is None: base = a
    k = addr2line.get(a - base, ('?', 0)); agg[k][0] += n; agg[k][1] += s
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total', tot, 'per warp-round', tot / nrounds / nwarps, 'static', len(data))
import glob, os
src = {os.path.basename(f): open(f).read().split('\n') for f in glob.glob('include/*') + glob.glob('stateline_b200/plugins/*.cu')}
def fn_of(f, l):
    if f not in src: return f
    L = src[f]
    for i in range(min(l - 1, len(L) - 1), -1, -1):
        m = re.match(r'^(?:__device__|SLM_FN|__global__|inline|static)[^;]*?(\w+)\(', L[i])
        if m: return f[6:14] + ':' + m.group(1)
    return f
reg = collections.defaultdict(lambda: [0, 0])
for k, v in agg.items():
    key = fn_of(*k); reg[key][0] += v[0]; reg[key][1] += v[1]
for k, v in sorted(reg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[6]) if len(sys.argv) > 6 else 20]:
    print('%-34s instr %5.1f%% (%6.0f /warp-round) samples %5.1f%%' % (k, 100 * v[0] / tot, v[0] / nrounds / nwarps, 100 * v[1] / ts))
