"""Kernel timeline of one captured training step (CUPTI activity records through torch.profiler).

    B=128 python tools/timeline.py [config] > gpurun_out/timeline_b128.txt

Prints, for the median of the profiled replays: the step's span, the time during which no kernel was
running at all (launch / dependency latency), the mean number of kernels in flight, per-stream busy time
and the launches in start order with their stream, start offset, duration and the gap to the previous
kernel's end on the same stream.  Tracing perturbs the timing slightly (CUPTI), so spans are compared with
the untraced step time printed first; nothing here is a bench value.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else 'cifar10-ac'
    B = int(os.environ.get('B', '128'))
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    run = bench.Run(config, B, os.environ.get('PREC', 'bf16'), dev, 0, 1)
    run.warm(5)
    eng, plan = run.eng, run.plan
    eng._feed(plan, run.feed(0), True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.run_resident(plan, True)
    e1.record()
    torch.cuda.synchronize()
    print('# untraced: %.1f us / step (back to back, L2 warm)' % (e0.elapsed_time(e1) * 1000 / 50))
    from torch.profiler import ProfilerActivity, profile
    nrep = 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(nrep):
            eng.run_resident(plan, True)
            torch.cuda.synchronize()
    class Ev:
        def __init__(self, k):
            self.name, self.lo, self.hi, self.stream = k.name(), k.start_ns() / 1e3, k.end_ns() / 1e3, k.device_resource_id()
    evs = [Ev(k) for k in prof.profiler.kineto_results.events()
           if k.device_type() == torch.autograd.DeviceType.CUDA and 'emcpy' not in k.name() and 'emset' not in k.name()]
    evs.sort(key=lambda e: e.lo)
    # split into replays: a gap > 30 us with nothing running separates them (synchronize in between)
    steps, cur, end = [], [], None
    for e in evs:
        if end is not None and e.lo - end > 30 and cur:
            steps.append(cur)
            cur = []
        cur.append(e)
        end = e.hi if end is None else max(end, e.hi)
    if cur:
        steps.append(cur)
    steps = [s for s in steps if len(s) > 20]
    spans = sorted((max(e.hi for e in s) - s[0].lo, i) for i, s in enumerate(steps))
    span, idx = spans[len(spans) // 2]
    s = steps[idx]
    t0 = s[0].lo
    print('# traced replays: %d, kernels in the median replay: %d, span %.1f us' % (len(steps), len(s), span))
    # union of busy intervals, mean concurrency
    iv = sorted((e.lo - t0, e.hi - t0) for e in s)
    busy, cur_lo, cur_hi = 0.0, iv[0][0], iv[0][1]
    for lo, hi in iv[1:]:
        if lo > cur_hi:
            busy += cur_hi - cur_lo
            cur_lo, cur_hi = lo, hi
        else:
            cur_hi = max(cur_hi, hi)
    busy += cur_hi - cur_lo
    total = sum(hi - lo for lo, hi in iv)
    print('# some kernel running: %.1f us (%.0f %%), idle: %.1f us, sum of kernel durations %.1f us, '
          'mean kernels in flight while busy %.2f' % (busy, 100 * busy / span, span - busy, total, total / busy))
    by_stream = {}
    for e in s:
        by_stream.setdefault(e.stream, []).append(e)
    for st, es in sorted(by_stream.items(), key=lambda kv: kv[1][0].lo):
        print('# stream %-4s n=%3d  busy %7.1f us  first %7.1f  last end %7.1f' % (
            st, len(es), sum(e.hi - e.lo for e in es), es[0].lo - t0, max(e.hi for e in es) - t0))
    last_end = {}
    print('# %-5s %9s %8s %8s  %s' % ('strm', 'start', 'dur', 'gap', 'kernel'))
    for e in s:
        st = e.stream
        lo, hi = e.lo - t0, e.hi - t0
        gap = lo - last_end[st] if st in last_end else float('nan')
        last_end[st] = hi
        print('  %-5s %9.1f %8.1f %8.1f  %s' % (st, lo, hi - lo, gap, e.name[:90]))


if __name__ == '__main__':
    main()
