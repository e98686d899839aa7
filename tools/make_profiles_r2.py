"""Round 2: turn the text exports of a gpurun evidence call (tools/ncu_export.sh: gpurun_out/r2_*_raw.csv, *_src.csv,
*_details.txt, r2_launches.csv) and the bench lines into the tracked summaries under profiles/."""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
sys.path.insert(0, ROOT)
from stateline_b200 import build

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'wait', 'short_scoreboard', 'mio_throttle', 'no_instruction', 'barrier', 'branch_resolving', 'lg_throttle',
         'math_pipe_throttle', 'not_selected', 'dispatch_stall')] + \
       ['smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
MUL = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}


def metrics(tag, out_md, title, command):
    rows = list(csv.reader(open(os.path.join(G, tag + '_raw.csv'))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    lines = ['# %s' % title, '', 'Command: `%s`' % command, '', 'Kernel: `%s`' % data[0][ki][:110], '',
             '| metric | unit | ' + ' | '.join('launch %d' % (i + 1) for i in range(len(data))) + ' |', '|---|---|' + '---|' * len(data)]
    out = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out[w] = (units[i], [r[i] for r in data])
            lines.append('| %s | %s | %s |' % (w, units[i], ' | '.join(r[i] for r in data)))
    open(os.path.join(P, out_md), 'w').write('\n'.join(lines) + '\n')
    f = lambda x: float(x.replace(',', ''))
    tr = [f(a) * MUL[out['dram__bytes_read.sum'][0]] + f(b) * MUL[out['dram__bytes_write.sum'][0]]
          for a, b in zip(out['dram__bytes_read.sum'][1], out['dram__bytes_write.sum'][1])]
    return out, tr


def per_line(tag, pat, top=14):
    """per-source-line split (tools/ncu_lines.py) against the disassembly of the in-tree plugin build"""
    so = build.TARGETS_TILE_SO if 'tile' in tag else build.TARGETS_SO
    tmp = '/tmp/_mk_r2'
    os.makedirs(tmp, exist_ok=True)
    subprocess.check_call('cd %s && rm -f *.cubin && cuobjdump -xelf all %s > /dev/null && nvdisasm --print-line-info $(ls *.cubin | head -1) > dis.txt' % (tmp, so), shell=True)
    return subprocess.check_output([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), os.path.join(tmp, 'dis.txt'),
                                    os.path.join(G, tag + '_src.csv'), pat, str(top)], text=True)


# ---- launch list ----
shutil.copy(os.path.join(G, 'r2_launches.csv'), os.path.join(P, 'r2_launches.csv'))
rows = [r for r in csv.reader(open(os.path.join(P, 'r2_launches.csv'))) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki].split('(')[0][:70], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
lines = ['Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv … python bench.py --steps 5 --warmup 3 --no-ess --no-cpu-baseline --no-side`',
         '(after the same command had exited 0 without ncu).  Times under ncu are serialised and cold: shares, not absolutes.', '',
         '| kernel | launches | avg us | total us | share |', '|---|---|---|---|---|']
for k, v in agg.items():
    lines.append('| `%s` | %d | %.1f | %.1f | %.1f%% |' % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
step = sum(sum(v) for k, v in agg.items() if 'pt_step_fast' in k)
tier = sum(sum(v) for k, v in agg.items() if 'k_tier_adapt' in k)
comp = sum(sum(v) for k, v in agg.items() if 'k_compact_rows' in k)
lines += ['', 'pt_step_fast share of step + tier + compaction kernels: %.1f%% (of step + tier alone: %.1f%%)' % (100 * step / (step + tier + comp), 100 * step / (step + tier))]
open(os.path.join(P, 'r2_launches_summary.md'), 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))

# ---- the dominant kernel ----
BC = 'python bench.py --steps 5 --warmup 3 --no-ess --no-cpu-baseline --no-side'
out, tr = metrics('r2_fast', 'r2_pt_step_fast_ncu_metrics.md', 'pt_step_fast_kernel, config #3, two launches of 10 rounds (ncu --set full)',
                  'ncu --set full --clock-control none --import-source on -k regex:pt_step_fast -s 5 -c 2 … ' + BC)
open(os.path.join(P, 'r2_pt_step_fast_ncu_metrics.md'), 'a').write('\n## Executed instructions and stall samples per source line\n\n```\n' + per_line('r2_fast', 'pt_step_fast_kernel.*GaussFact', 16) + '```\n')
shutil.copy(os.path.join(G, 'r2_fast_details.txt'), os.path.join(P, 'r2_pt_step_fast_ncu_details.txt'))
digest = open(build.TARGETS_SO + '.srchash').read().strip()[:16]
json.dump({'kernel': 'pt_step_fast_kernel<GaussFact,20,128>',
           'source': 'profiles/r2_pt_step_fast_ncu_metrics.md (ncu --set full, 2 launches of 10 rounds, config #3)',
           'dram_bytes_per_launch': sum(tr) / len(tr), 'per_launch': tr, 'plugin_digest': digest},
          open(os.path.join(P, 'pt_step_traffic.json'), 'w'), indent=1)
print('traffic', tr, 'digest', digest)

# ---- the opt-in tile kernel ----
out, tr = metrics('r2_tile', 'r2_pt_step_tile_ncu_metrics.md', 'pt_step_tile_kernel (opt-in, -DSLGPU_TILE_KERNEL), config #3, one launch of 10 rounds',
                  'ncu --set full --clock-control none --import-source on -k regex:pt_step_tile -s 5 -c 1 … python tools/probe.py --workload config3 --warm 60 --rounds 30 --plugin stateline_b200/libslgpu_targets_tile.so')
print('tile traffic', tr)
sass = subprocess.check_output('cuobjdump -sass %s | grep -E "Function|UBLKCP|SYNCS|UTMA" | grep -B1 -E "UBLKCP|SYNCS" | head -60' % build.TARGETS_TILE_SO, shell=True, text=True)
open(os.path.join(P, 'r2_tile_kernel_sass.txt'), 'w').write(
    'cuobjdump -sass stateline_b200/libslgpu_targets_tile.so | grep -E "Function|UBLKCP|SYNCS": the bulk asynchronous copies (UBLKCP.S.G = global -> shared,\n'
    'UBLKCP.G.S = shared -> global) and the mbarrier operations (SYNCS) of pt_step_tile_kernel\n\n' + sass)

# ---- the warp-per-chain kernel ----
md = ['# pt_step_warp_kernel on configs #4 and #5, steady state (200 rounds of warm-up), one launch of 10 rounds each', '']
for c, pat in ((4, 'pt_step_warp_kernel.*GaussCorr.*Li4E'), (5, 'pt_step_warp_kernel.*Rosenbrock.*Li2E')):
    out, tr = metrics('r2_w%d' % c, '_tmp.md', 'config #%d' % c,
                      'ncu --set full --clock-control none --import-source on -k regex:pt_step_warp -s 21 -c 1 … python tools/probe.py --workload config%d --warm 200 --rounds 30' % c)
    md += open(os.path.join(P, '_tmp.md')).read().replace('# config', '## config').split('\n')
    md += ['', 'DRAM traffic per launch: %.3g B' % tr[0], '', '### per source line', '', '```'] + per_line('r2_w%d' % c, pat, 14).split('\n') + ['```', '']
os.remove(os.path.join(P, '_tmp.md'))
open(os.path.join(P, 'r2_warp_kernel_ncu.md'), 'w').write('\n'.join(md) + '\n')

# ---- the re-Cholesky kernel of the opt-in welford mode ----
if os.path.exists(os.path.join(G, 'r2_rechol_raw.csv')):
    out, tr = metrics('r2_rechol', 'r2_rechol_ncu_metrics.md', 'rechol_kernel<20> ("gpu": {"shaper": "welford"}), config #3, one adapt point',
                      'ncu --set full --clock-control none --import-source on -k regex:rechol -s 3 -c 1 … python tools/probe.py --workload config3 --warm 60 --rounds 30 --shaper welford')
    print('rechol traffic', tr)
print('done')
