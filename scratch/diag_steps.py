"""Per-step, per-tensor comparison of net.train.run against the fp64 oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')]
import numpy as np, torch
import make_golden as mg
from oracle.torch_ref import OracleNet
from util import batch, record_of

name = sys.argv[1] if len(sys.argv) > 1 else 'sr'
net2 = mg.build(name).configure(precision='fp32')
o2 = OracleNet(record_of(mg.build(name)), torch.float64)
e2 = net2._get_engine()
print('hypers', vars(net2.hypers))
for t in range(3):
    xb, yb = batch(16, seed=10 + t)
    f = {net2.x0: xb, net2.y: yb, net2.mode: 'tr', net2.λ_lrn: 0.05 / 2 ** t}
    if net2.dynamic:
        f[net2.τ] = 1.0 / 2 ** (t / 2)
    net2.train.run(f)
    out = o2.train_step(xb, yb, lr=0.05 / 2 ** t, mu=0.9, tau=1.0 / 2 ** (t / 2))
    th = [e2._buf(p).double().cpu().numpy() for p in e2.tparams]
    print('step', t, 'c_tot oracle', float(out.c_tot))
    for i, ((pp, rr, kk, tt), a) in enumerate(zip(o2.trainable, th)):
        b = tt.detach().numpy()
        d = np.abs(a.reshape(-1) - b.reshape(-1)).max()
        print('  %2d %-8s %-7s %-6s shape %-16s |ours| %.6e |orc| %.6e maxdiff %.3e' % (
            i, pp, rr, kk, tuple(b.shape), np.linalg.norm(a), np.linalg.norm(b), d))
