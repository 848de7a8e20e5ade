"""Turns the files a tools/gpu_final.sh run left in gpurun_out/ into the committed summaries under profiles/."""
import collections
import csv
import json
import re
import shutil
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = [r for r in csv.reader(open('gpurun_out/launches_ncu_%s.csv' % R)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, tot = collections.OrderedDict(), 0.0
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    ns = v * 1e3 if r[ui] == 'us' else (v * 1e6 if r[ui] == 'ms' else v)
    name = re.sub(r'\(.*', '', r[ki]).replace('void mf::', '').replace('void ', '')
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
lines = ['# ncu launch list (bench.py --steps 1 --warmup 1, `--metrics gpu__time_duration.sum --clock-control none`), ' + R,
         '# cold-cache serialised per-launch times: compare SHARES. total %.3f ms over %d launches' % (tot / 1e6, sum(a[0] for a in agg.values())),
         '# igemm2_kernel<BLOCK_N, MODE, NPW>: MODE 0 = cp.async gather, 1 = DCN gather, 2 = im2col TMA',
         '%-62s %6s %10s %7s' % ('kernel', 'n', 'sum_us', 'share')]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append('%-62s %6d %10.1f %6.1f%%' % (k[:62], a[0], a[1] / 1e3, 100 * a[1] / tot))
open('profiles/ncu_launch_list_%s.txt' % R, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:16]))

rows = list(csv.reader(open('gpurun_out/prof_raw_%s.csv' % R)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']
out = ['# ncu --set full --clock-control none, selected launches of the B=8 inference plan (tools/profile_kernels.py), ' + R]
traffic = {}


def val(r, k):
    i = hdr.index(k)
    return float(r[i].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[i], 1)


for n, r in enumerate(rows[2:]):
    kn = r[hdr.index('Kernel Name')]
    grid = r[hdr.index('launch__grid_size')]
    out.append('\n## launch %d: %s   grid %s' % (n, kn[:100], grid))
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            out.append('   %-92s %s %s' % (k, r[i], units[i]))
    if n == 0:
        key = 'head_fused_dram_bytes_per_launch' if 'head_fused' in kn else 'head_conv_dram_bytes_per_launch'
        traffic[key] = val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
open('profiles/ncu_full_summary_%s.txt' % R, 'w').write('\n'.join(out) + '\n')
json.dump(traffic, open('profiles/roofline_traffic.json', 'w'), indent=1)
for f in ('bench_%s.json', 'bench_ref_%s.json', 'launches_events_%s.json', 'clocks_%s.csv', 'pytest_gpu_%s.log', 'smoke_%s.log'):
    try:
        shutil.copy('gpurun_out/' + f % R, 'profiles/' + f % R)
    except FileNotFoundError:
        pass
print(traffic)
