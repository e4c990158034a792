"""Turn what tools/profile_round.sh left in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU):

    python tools/collect_profiles.py

* copies the bench lines, per-kernel tables, timelines, critical paths, test / smoke logs, clocks;
* r02_ncu_launch_summary.csv: per-kernel totals and shares of the serialised launch list;
* r02_ncu_full_<what>.csv: the metrics the design argues with, from every `ncu --set full` capture;
* r02_traffic.json: measured DRAM bytes per launch of the dominant kernels, keyed the way bench.py looks them up;
* r02_sass_mnemonics.txt: tcgen05 / TMA evidence from the built library.
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')

COPY = ['r02_bench_default.json', 'r02_bench_per_kernel.txt', 'r02_bench_b128.json', 'r02_bench_per_kernel_b128.txt',
        'r02_bench_reference_arm.json', 'r02_timeline_b128.txt', 'r02_timeline_b4096.txt', 'r02_critpath_b128.txt',
        'r02_critpath_b4096.txt', 'r02_gpu_tests.log', 'r02_smoke.log', 'r02_nvidia_smi.txt', 'r02_ncu_launch_list.csv',
        'r02_dp2_bench.json', 'r02_dp2_tests.log', 'r02_dp8_bench.json', 'r02_mb_chain.txt']

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
           'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'sm__cycles_elapsed.max',
           'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warp_latency_issue_stalled_barrier.pct']


def ncu_raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarise(rep, dst):
    head, units, rows = ncu_raw(rep)
    idx = {n: i for i, n in enumerate(head)}
    cols = ['Kernel Name'] + [m for m in METRICS if m in idx]
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows:
            w.writerow([r[idx[c]] for c in cols])
    return [(r[idx['Kernel Name']], {c: r[idx[c]] for c in cols}) for r in rows]


def to_bytes(v, unit):
    v = float(v)
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main():
    os.makedirs(PROF, exist_ok=True)
    for f in COPY:
        if os.path.exists(os.path.join(OUT, f)):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f))
    # launch list -> per-kernel totals
    ll = os.path.join(OUT, 'r02_ncu_launch_list.csv')
    if os.path.exists(ll):
        text = open(ll).read()
        text = text[text.index('"ID"'):]
        rows = list(csv.DictReader(io.StringIO(text)))
        tot = OrderedDict()
        for r in rows:
            if r.get('Metric Name') != 'gpu__time_duration.sum':
                continue
            name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('(anonymous namespace)::', '')
            v = float(r['Metric Value']) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r['Metric Unit'], 1.0)
            d = tot.setdefault(name, [0, 0.0])
            d[0] += 1; d[1] += v
        s = sum(v for _, v in tot.values())
        with open(os.path.join(PROF, 'r02_ncu_launch_summary.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['kernel', 'launches', 'total_us', 'share'])
            for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, n, '%.1f' % v, '%.4f' % (v / s)])
    # full captures
    traffic = {}
    reps = {'r02_prof_conv_fwd': ('r02_ncu_full_conv_fwd_h32.csv', 'B4096', 'conv_fwd H32 K16+0 N16'),
            'r02_prof_conv_dgrad': ('r02_ncu_full_conv_dgrad_h32.csv', 'B4096', 'conv_dgrad H32 K16 N16+0 +bnred'),
            'r02_prof_wgrad_h32': ('r02_ncu_full_wgrad_h32.csv', 'B4096', 'conv_wgrad H32 K16+0 N16'),
            'r02_prof_wgrad_h8': ('r02_ncu_full_wgrad_h8.csv', 'B4096', 'conv_wgrad H8 K64+0 N64'),
            'r02_prof_bn': ('r02_ncu_full_bn_h32.csv', 'B4096', None),
            'r02_prof_h4_b128': ('r02_ncu_conv_h4_b128.csv', None, None),
            'r02_prof_h4_b4096': ('r02_ncu_conv_h4_b4096.csv', None, None),
            'r02_prof_router_b128': ('r02_ncu_small_kernels_b128.csv', None, None),
            'r02_prof_k32n64': ('r02_ncu_full_conv_k32n64.csv', None, None)}
    for rep, (dst, bkey, lname) in reps.items():
        path = os.path.join(OUT, rep + '.ncu-rep')
        if not os.path.exists(path):
            continue
        rows = summarise(path, os.path.join(PROF, dst))
        head, units, _ = ncu_raw(path)
        u = {n: units[i] for i, n in enumerate(head)}
        for name, m in rows:
            if 'dram__bytes_read.sum' not in m:
                continue
            b = to_bytes(m['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + to_bytes(m['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
            key = lname
            if rep == 'r02_prof_bn':
                key = 'bn_bwd H32 C16' if 'bwd_v2' in name or 'pool_bwd' in name else 'bn_fwd H32 C16' if 'pool_fwd' in name else None
            if bkey and key:
                traffic.setdefault(bkey, {})[key] = {
                    'dram_bytes': b, 'source': 'profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of this launch)' % dst}
    if traffic:
        tpath = os.path.join(PROF, 'r02_traffic.json')
        merged = json.load(open(tpath)) if os.path.exists(tpath) else {}       # captures of earlier calls stay
        for bkey, d in traffic.items():
            merged.setdefault(bkey, {}).update(d)
        merged.get('B4096', {}).pop('bn_fwd', None)
        json.dump(merged, open(tpath, 'w'), indent=1)
    raw = os.path.join(OUT, 'r02_ncu_routing_kernels_raw.csv')
    if os.path.exists(raw):
        rows = list(csv.reader(open(raw)))
        head, units, body = rows[0], rows[1], rows[2:]
        idx = {n: i for i, n in enumerate(head)}
        cols = ['Kernel Name'] + [m for m in METRICS if m in idx]
        with open(os.path.join(PROF, 'r02_ncu_routing_kernels.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(cols); w.writerow([units[idx[c]] for c in cols])
            for r in body:
                w.writerow([r[idx[c]] for c in cols])
    # SASS evidence
    lib = os.path.join(ROOT, 'multipath-nn_b200', 'lib', 'libmpnn_sm100.so')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    counts = {}
    for mn in ('UTCHMMA', 'UTCBAR', 'LDTM', 'UBLKCP', 'SYNCS', 'UTCATOMSWS', 'ACQBULK', 'UCGABAR', 'REDG', 'RED.E'):
        counts[mn] = len(re.findall(r'\b' + re.escape(mn), sass))
    with open(os.path.join(PROF, 'r02_sass_mnemonics.txt'), 'w') as f:
        f.write('cuobjdump -sass multipath-nn_b200/lib/libmpnn_sm100.so | grep -c <mnemonic>\n')
        for k, v in counts.items():
            f.write('%-12s %d\n' % (k, v))
    print('profiles/:', sorted(os.listdir(PROF)))


if __name__ == '__main__':
    sys.exit(main())
