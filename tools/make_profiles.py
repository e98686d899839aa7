"""Turn the ncu outputs of a gpurun call (gpurun_out/launches_r1.csv, prof_r1.ncu-rep, bench json) into the
tracked summaries under profiles/ (launch list, metric table, traffic json)."""
import collections, csv, json, subprocess, sys

bench = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench_r1_v6.json'
subprocess.check_call('cp gpurun_out/launches_r1.csv profiles/r1_launches.csv && cp %s profiles/r1_bench.json' % bench, shell=True)
subprocess.check_call('ncu -i gpurun_out/prof_r1.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/raw_r1.csv', shell=True)
subprocess.check_call('ncu -i gpurun_out/prof_r1.ncu-rep --page details 2>/dev/null > profiles/r1_pt_step_ncu_details.txt', shell=True)
rows = [r for r in csv.reader(open('profiles/r1_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki].split('(')[0][:70], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
lines = ['| kernel | launches | avg us | total us | share |', '|---|---|---|---|---|']
for k, v in agg.items():
    lines.append('| `%s` | %d | %.1f | %.1f | %.1f%% |' % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
open('profiles/r1_launches_summary.md', 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
rows = list(csv.reader(open('gpurun_out/raw_r1.csv'))); hdr = rows[0]; units = rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'wait', 'short_scoreboard', 'mio_throttle', 'no_instruction', 'barrier', 'branch_resolving', 'lg_throttle', 'math_pipe_throttle')] + \
       ['smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size']
out = {}
lines = ['| metric | unit | launch 1 | launch 2 |', '|---|---|---|---|']
for w in want:
    if w in hdr:
        i = hdr.index(w); vals = [r[i] for r in rows[2:]]; out[w] = vals
        lines.append('| %s | %s | %s |' % (w, units[i], ' | '.join(vals)))
open('profiles/r1_pt_step_ncu_metrics.md', 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
f = lambda x: float(x.replace(',', ''))
mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
ui = units[hdr.index('dram__bytes_read.sum')]; uw = units[hdr.index('dram__bytes_write.sum')]
tr = [f(a) * mul[ui] + f(b) * mul[uw] for a, b in zip(out['dram__bytes_read.sum'], out['dram__bytes_write.sum'])]
json.dump({'kernel': 'pt_step_fast_kernel<GaussFact,20,128>',
           'source': 'profiles/r1_pt_step_ncu_metrics.md (ncu --set full, 2 launches of 10 rounds, config #3)',
           'dram_bytes_per_launch': sum(tr) / len(tr), 'per_launch': tr}, open('profiles/pt_step_traffic.json', 'w'), indent=1)
b = json.load(open('profiles/r1_bench.json'))
print(tr, {k: b[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks']}, b['e2e']['value'])
print(b['roofline']['frac'], b['roofline']['achieved'], b['roofline']['launch_ms'], b['roofline']['share_of_step'], b['roofline']['traffic'],
      b['roofline']['fp64']['frac'], b['roofline']['fp64']['achieved'])
print(b['cpu_baseline']['value'], b['cpu_baseline']['inline_1_thread']['value'], b.get('ess'))
