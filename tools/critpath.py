"""Critical path of the step's op DAG (lanes + explicit deps) with per-op eager CUDA-event times."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from lib import layer_types
import arch_and_hypers as ah

B = int(os.environ.get('B', 128))
layer_types.seed(0)
net = ah.ac_chain(k_cpt=4e-9)((32, 32, 3), (10,)).configure(precision='bf16')
eng = net._get_engine()
rng = np.random.default_rng(0)
x0 = rng.random((B, 32, 32, 3)).astype(np.float32); y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
feed = {net.x0: x0, net.y: y, net.τ: 1.0, net.mode: 'tr'}
for _ in range(3):
    net.train.run(feed)
torch.cuda.synchronize()
plan = eng._plan(B, True, True)
lists = [('pack', plan.pack_ops), ('fwd', plan.fwd_ops), ('bwd', plan.bwd_ops), ('opt', plan.opt_ops)]
called = []
L = eng.L
saved = {}
for name in L.protos:
    short = name[5:]
    try: fn = getattr(L, short)
    except AttributeError: continue
    saved[short] = fn
    L.__dict__[short] = (lambda fn, short: (lambda *a: (called.append(short), fn(*a))[1]))(fn, short)
times = {}
for rep in range(3):
    eng.grad.zero_()
    eng.stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    evs = []
    for lname, ops in lists:
        for op in ops:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            del called[:]
            a.record(); op(); b.record()
            if not hasattr(op, 'kind'): op.kind = called[0] if called else 'misc'
            evs.append((op, a, b))
    torch.cuda.synchronize()
    for op, a, b in evs: times[id(op)] = a.elapsed_time(b) * 1e3
tot_path = 0.0
for lname, ops in lists:
    end, lane_end, pred = {}, {}, {}
    for op in ops:
        lane = getattr(op, 'lane', 0)
        cands = [(lane_end.get(lane, (0.0, None)))] + [(end[id(d)], d) for d in getattr(op, 'deps', ()) if id(d) in end]
        st, p = max(cands, key=lambda c: c[0])
        end[id(op)] = st + times[id(op)]; pred[id(op)] = p
        lane_end[lane] = (end[id(op)], op)
    last = max(ops, key=lambda o: end[id(o)]) if ops else None
    if last is None: continue
    path = []
    o = last
    while o is not None:
        path.append(o); o = pred[id(o)]
    path.reverse()
    print('== %s: %d ops, serial %.0f us, critical path %.0f us over %d ops' % (
        lname, len(ops), sum(times[id(o)] for o in ops), end[id(last)], len(path)))
    tot_path += end[id(last)]
    for o in path:
        print('   lane %d %-24s %-20s %6.1f us' % (getattr(o, 'lane', 0), o.kind, getattr(o, 'desc', ''), times[id(o)]))
print('critical path total %.0f us (eager per-op times include ~3 us of launch overhead each)' % tot_path)
