#!/usr/bin/env python3
"""Benchmark of the hot path: one training step of the `cifar10-ac` net
(ac_chain, /root/reference/scripts/arch_and_hypers.py:76-139, experiment table
scripts/train-nets:84-87) on synthetic CIFAR-10-shaped data.

  python bench.py --gpus N --steps K --warmup W [--batch B] [--precision bf16|fp32]
  python bench.py --impl reference ...      # CPU oracle arm (reference semantics on PyTorch-CPU)

One JSON line on rank 0 (see the driver contract).  `value` = images/s with
the batch already resident in HBM (device timed, per-step CUDA events, L2
flushed between steps); `e2e` = images/s through net.train.run(feed) with
pinned HOST batches (H2D inside the timed region, loss vector read back).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

K_CPT = 4e-9            # k_cpts[3] of arch_and_hypers.py:15
TRAIN_FLOP_PER_IMG = 123.02e6      # BASELINE.md section 3 (ac_chain, dense as in the reference)


def lam(t):
    return 0.1 / 2 ** (t / 10000)


def tau_ds(t):
    return 1 / 2 ** (t / 20000)


def make_net(seed=0):
    from lib import layer_types
    import arch_and_hypers as ah
    layer_types.seed(seed)
    return ah.ac_chain(k_cpt=K_CPT)((32, 32, 3), (10,))


def synth(B, n, seed):
    rng = np.random.default_rng(seed)
    xs = [torch.from_numpy(rng.random((B, 32, 32, 3), dtype=np.float32)) for _ in range(n)]
    ys = [torch.from_numpy(np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]) for _ in range(n)]
    return xs, ys


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class Clocks:
    """nvidia-smi sampler (recipe's clocks line) running during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.rows, self.proc, self.gpu, self.mark = [], None, gpu, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        # samples of the timed regions; a very short run (few steps) may have none, then the samples taken
        # under the warm-up load stand in and the line says so
        rows = self.rows[self.mark:]
        window = 'timed region'
        if not rows:
            rows, window = list(self.rows), 'warm-up (timed region shorter than the sampling interval)'
        sm = [float(r[1]) for r in rows if len(r) > 2 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'window': window}


def cpu_oracle_rate(B, budget_s=20.0, max_steps=50):
    """Reference semantics restated on PyTorch-CPU (oracle/torch_ref.py): full
    train steps (forward 'tr', autograd backward, TALR + momentum) at batch B."""
    import copy
    from lib import serdes
    from oracle.torch_ref import OracleNet
    net = make_net()
    o = OracleNet(copy.deepcopy(serdes.encode_net(net)), torch.float32)
    xs, ys = synth(B, 2, 1)
    o.train_step(xs[0].numpy(), ys[0].numpy(), lr=lam(0), tau=tau_ds(0))        # warm-up
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < budget_s):
        o.train_step(xs[n % 2].numpy(), ys[n % 2].numpy(), lr=lam(n), tau=tau_ds(n))
        n += 1
    dt = time.perf_counter() - t0
    return n * B / dt, n, dt


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  TensorFlow <= 0.12 cannot be installed
    offline, so this times the oracle port of the reference on the host cores."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch
    steps = max(1, args.steps)
    import copy
    from lib import serdes
    from oracle.torch_ref import OracleNet
    net = make_net()
    o = OracleNet(copy.deepcopy(serdes.encode_net(net)), torch.float32)
    sample_B = min(B, 256)                     # bounded sample of the per-GPU batch
    xs, ys = synth(sample_B, 2, 1)
    for w in range(min(args.warmup, 2)):
        o.train_step(xs[0].numpy(), ys[0].numpy(), lr=lam(w), tau=tau_ds(w))
    k = min(steps, 10)
    t0 = time.perf_counter()
    for t in range(k):
        o.train_step(xs[t % 2].numpy(), ys[t % 2].numpy(), lr=lam(t), tau=tau_ds(t))
    dt = time.perf_counter() - t0
    v = k * sample_B / dt
    line = {
        'impl': 'reference', 'metric': 'train images/sec', 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': k, 'warmup': min(args.warmup, 2), 'ms_per_step': 1e3 * dt / k, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'cifar10-ac: ac_chain(k_cpt=4e-9) train step, 32x32x3, 10 classes',
                   'batch_per_gpu': B, 'sample_batch': sample_B},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': '%d train steps at batch %d (reference semantics restated on PyTorch-CPU; '
                                   'TF<=0.12 not installable offline)' % (k, sample_B)},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=int(os.environ.get('MPNN_BENCH_BATCH', 4096)),
                    help='examples per GPU per step (reference trains at 128; see DESIGN.md)')
    ap.add_argument('--precision', default=os.environ.get('MPNN_PRECISION', 'bf16'))
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--profile', action='store_true', help='print per-kernel-kind time shares')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    B = args.batch
    net = make_net().configure(precision=args.precision, graphs=not args.no_graphs, dist=world > 1)
    eng = net._get_engine()
    L = eng.L
    xs, ys = synth(B, 4, 100 + rank)
    xs = [x.pin_memory() for x in xs]
    ys = [y.pin_memory() for y in ys]
    plan = eng._plan(B, True, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def feed(t):
        return {net.x0: xs[t % 4], net.y: ys[t % 4], net.mode: 'tr', net.λ_lrn: lam(t), net.τ: tau_ds(t)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def note(msg):
        if os.environ.get('MPNN_BENCH_VERBOSE'):
            print('[bench rank %d] %s' % (rank, msg), file=sys.stderr, flush=True)

    # ---------------- warm-up (also captures the CUDA graphs) --------------- #
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()                                 # nvidia-smi needs ~0.2 s before its first sample
    note('warm-up')
    for t in range(args.warmup):
        net.train.run(feed(t))
    if rank == 0:
        # wait for the sampler WITHOUT issuing work: a training step contains the all-reduce, so every
        # rank must run exactly the same number of them
        torch.cuda.synchronize(dev)
        t_wait = time.perf_counter()
        while not clocks.rows and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.01)
        clocks.mark = len(clocks.rows)                 # samples from here on fall inside the timed regions
    barrier()
    note('warm-up done')

    # ---------------- device-resident timing -> value ---------------------- #
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    eng._feed(plan, feed(args.warmup), True)
    barrier()
    l0 = L.launches
    for t in range(args.steps):
        flush.zero_()                                  # evict L2 between timed iterations
        ev[t][0].record()
        eng.run_resident(plan, True)
        ev[t][1].record()
    barrier()
    launches = (plan.graph_launches * args.steps) if eng.use_graphs else (L.launches - l0)
    note('device timing done')
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms = float(tt.item())

    # ---------------- end-to-end timing through the public API -> e2e ------- #
    host_loss = torch.empty(B, dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for t in range(args.steps):
        net.train.run(feed(args.warmup + t))           # pinned host -> device copies inside
        host_loss.copy_(plan.c_data, non_blocking=True)  # per-example objective of the step
    torch.cuda.synchronize(dev)
    loss = float(host_loss.mean())
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s = float(tt.item())
    clk = clocks.stop() if rank == 0 else None
    note('e2e done')

    # ---------------- per-kernel profile (eager, CUDA events per launch) ---- #
    prof = {}
    if rank == 0:
        ops = plan.pack_ops + plan.fwd_ops + plan.bwd_ops + plan.opt_ops
        # untagged launches are labelled by the C-ABI entry point they call
        called = []
        saved = {}
        for name in L.protos:
            short = name[5:]
            try:
                fn = getattr(L, short)
            except AttributeError:
                continue
            saved[short] = fn
            L.__dict__[short] = (lambda fn, short: (lambda *a: (called.append(short), fn(*a))[1]))(fn, short)
        for rep in range(3):
            evs = []
            eng.grad.zero_()
            eng.stream = __import__('ctypes').c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for op in ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                del called[:]
                a.record(); op(); b.record()
                if not hasattr(op, 'kind'):
                    op.kind = called[0] if called else 'misc'
                evs.append((op, a, b))
            torch.cuda.synchronize(dev)
            if rep == 2:
                L.__dict__.update(saved)
                for op, a, b in evs:
                    k = getattr(op, 'kind', 'misc')
                    d = prof.setdefault(k, {'ms': 0.0, 'flops': 0.0, 'bytes': 0.0, 'n': 0})
                    d['ms'] += a.elapsed_time(b); d['flops'] += getattr(op, 'flops', 0.0)
                    d['bytes'] += getattr(op, 'nbytes', 0.0); d['n'] += 1
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    tot_ms = sum(d['ms'] for d in prof.values())
    top = max(prof, key=lambda k: prof[k]['ms'])
    d = prof[top]
    # binding ceiling of the dominant kernel: the larger of its HBM time (algorithmic bytes /
    # measured copy bandwidth) and its tensor time (flops / measured cuBLAS bf16 rate).  The thin
    # convolutions of this net (16..32 channels; 72..144 flop/B) sit below the ridge (~200 flop/B).
    t_hbm = d['bytes'] / (pk['hbm_gbs'] * 1e9)
    t_tc = d['flops'] / (pk['bf16_tflops'] * 1e12)
    ach_tf = d['flops'] / (d['ms'] * 1e-3) / 1e12
    ach_gb = d['bytes'] / (d['ms'] * 1e-3) / 1e9
    if t_tc > t_hbm:
        roof = {'bound': 'tensor', 'achieved': ach_tf, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': ach_tf / pk['bf16_tflops'], 'traffic': None,
                'other_ceiling': {'bound': 'hbm', 'achieved': ach_gb, 'frac': ach_gb / pk['hbm_gbs']}}
    else:
        roof = {'bound': 'hbm', 'achieved': ach_gb, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                'frac': ach_gb / pk['hbm_gbs'], 'traffic': None,
                'other_ceiling': {'bound': 'tensor', 'achieved': ach_tf, 'frac': ach_tf / pk['bf16_tflops']}}
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json'))).get(top)
        if tr:
            roof['traffic'] = tr['dram_bytes_per_launch']
            roof['traffic_of'] = '%s: %d algorithmic bytes (%s)' % (tr['launch'], tr['algorithmic_bytes_per_launch'], tr['source'])
    except Exception:
        pass
    roof.update({'kernel': top, 'launches_per_step': d['n'], 'share_of_step': d['ms'] / tot_ms,
                 'peak_source': pk_src + ' (burst; kernels timed one by one with CUDA events)',
                 'per_launch_avg_ms': d['ms'] / d['n']})
    shares = {k: round(v['ms'] / tot_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
    if args.profile:
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            extra = ''
            if v['flops']:
                extra = '%.1f TFLOP/s' % (v['flops'] / (v['ms'] * 1e-3) / 1e12)
            elif v['bytes']:
                extra = '%.0f GB/s' % (v['bytes'] / (v['ms'] * 1e-3) / 1e9)
            print('# %-14s n=%3d %8.3f ms  %5.1f%%  %s' % (k, v['n'], v['ms'], 100 * v['ms'] / tot_ms, extra),
                  file=sys.stderr)
        for op, a, b in evs:
            if getattr(op, 'desc', ''):
                ms = a.elapsed_time(b)
                print('#   %-11s %-18s %7.3f ms %7.1f TFLOP/s %7.0f GB/s' % (
                    op.kind, op.desc, ms, op.flops / ms / 1e9, op.nbytes / ms / 1e6), file=sys.stderr)

    torch.set_num_threads(os.cpu_count() or 1)
    cpu_B = min(B, 256)
    cpu_v, cpu_n, cpu_dt = (0.0, 0, 0.0) if os.environ.get('MPNN_BENCH_NO_CPU') else cpu_oracle_rate(cpu_B, budget_s=15.0)
    value = world * B * args.steps / (dev_ms * 1e-3)
    e2e = world * B * args.steps / e2e_s
    h2d = B * (32 * 32 * 3 + 10) * 4 + 8 * 4
    line = {
        'metric': 'train images/sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': 'cifar10-ac: ac_chain(k_cpt=4e-9) train step (fwd tr + bwd + TALR/momentum), '
                               '32x32x3, 10 classes, dense as in the reference',
                   'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': 'dp%d' % world,
                   'l2': 'flushed between timed steps (256 MiB memset outside the event pairs)',
                   'cuda_graph': bool(eng.use_graphs), 'conv_impl': 'tcgen05' if eng.impl == 1 else 'simt',
                   'wgrad_impl': 'tcgen05' if eng.impl_w == 1 else 'simt'},
        'e2e': {'value': e2e, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': B * 4,
                'timing': 'wall clock around net.train.run(feed) with pinned host batches', 'loss': loss},
        'gpu_launches': int(launches),
        'tensor_frac_of_step': value / world * TRAIN_FLOP_PER_IMG / (pk['bf16_tflops_sustained'] * 1e12),
        # whole step against the HBM ceiling: algorithmic bytes of every launch (op tags) / measured copy bandwidth
        'hbm_frac_of_step': (sum(getattr(op, 'nbytes', 0.0) for op in ops) / (pk['hbm_gbs'] * 1e9)) / (dev_ms / args.steps * 1e-3),
        'roofline': roof, 'kernel_time_shares': shares,
        'cpu_baseline': {'value': cpu_v, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': '%d train steps at batch %d in %.1f s (reference semantics restated on '
                                   'PyTorch-CPU; TF<=0.12 not installable offline)' % (cpu_n, cpu_B, cpu_dt)},
        'clocks': clk}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
