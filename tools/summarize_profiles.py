"""Turn gpurun_out ncu artefacts into the committed summaries under profiles/ (development aid).

    python tools/summarize_profiles.py <round-tag> <launches.csv> <name>=<prof.ncu-rep | raw-page export .csv> [...]
"""
import collections
import csv
import json
import subprocess
import sys

tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]

# ---- launch list: per-kernel totals and shares -------------------------------
lines = [l for l in open(launches) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
# the timed region of `bench.py --steps 2`: the two segments (each opens with the time-dilation kernel of its
# boundary, after the 3 us halo wrap) before the FP64 probe
names = [r['Kernel Name'].split('(')[0] for r in rows]
end = names.index('fp64_rate_kernel') if 'fp64_rate_kernel' in names else len(names)
starts = [i for i, k in enumerate(names[:end]) if 'dilation_fused_kernel' in k]
rows = rows[starts[-2] if len(starts) >= 2 else 0:end]
agg = collections.OrderedDict()
for row in rows:
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6}[u]
    k = row['Kernel Name'].split('(')[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open('profiles/%s_launches_summary.txt' % tag, 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-fast --no-djds\n')
    f.write('# launches of the timed region (the two timed segments, cut out of the full launch list); times are cold-cache and serialised: compare SHARES\n')
    f.write('%-58s %7s %12s %10s %7s\n' % ('kernel', 'n', 'total_ms', 'avg_ms', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-58s %7d %12.3f %10.4f %6.1f%%\n' % (k[:58], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    f.write('%-58s %7d %12.3f\n' % ('TOTAL', sum(v[0] for v in agg.values()), tot))

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']

traffic, fp64 = {}, {}
for spec in reps:
    name, rep = spec.split('=')
    raw = open(rep).read() if rep.endswith('.csv') else \
        subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    with open('profiles/%s_%s_ncu.txt' % (tag, name), 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on : bench.py, N=2^24, m=32; one column per captured launch\n')
        f.write('# kernels: %s\n' % ' | '.join(r[ki] for r in data))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write('%-92s %-14s %s\n' % (w, units[i], ' | '.join(r[i] for r in data)))
    def val(r, w):
        i = hdr.index(w)
        x = float(r[i].replace(',', ''))
        return x * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[units[i]]
    for r in data:
        k = r[ki].split('(')[0].replace('void ', '').replace('fds::', '')
        traffic.setdefault(k, []).append(val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum'))
        fp64.setdefault(k, []).append(float(r[hdr.index('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')]))
out = {k: {'dram_bytes_per_launch': sum(v) / len(v), 'fp64_pipe_busy_pct': sum(fp64[k]) / len(fp64[k])}
       for k, v in traffic.items()}
if out:
    # what bench.py quotes as roofline.traffic: keyed by math mode, tied to the kernel sources by their hash
    sys.path.insert(0, '.')
    import bench
    sha = bench.step_sources_sha()
    doc = {'source': 'profiles/%s_*_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)' % tag,
           'workload': 'N=2^24 sites, m=32 (34 trajectories)', 'all_kernels': out}
    for k, v in out.items():
        if 'l96_rk4_stream' in k:
            mode = 'fast' if 'FastMath' in k else 'exact'
            doc[mode] = dict(v, kernel=k, src_sha16=sha)
    json.dump(doc, open('profiles/step_kernel_traffic.json', 'w'), indent=1)
print('ok', list(out))
