"""Is a large bf16 gradient deviation kernel error or sensitivity?  Compare two bf16-quantised ORACLES
that differ only in the accumulation dtype (fp64 vs fp32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from oracle.torch_ref import OracleNet
from util import batch, randomize_routers, record_of, tiny_net
for hy in (dict(k_cpt=4e-9), dict(k_cpt=4e-9, n_cls=5), dict(k_cpt=4e-9, x0_shape=(16, 16, 1), n_cls=5)):
    net = randomize_routers(tiny_net('ac', seed=0, **hy))
    rec = record_of(net)
    x0, y = batch(24, x0_shape=hy.get('x0_shape', (16, 16, 3)), n_cls=hy.get('n_cls', 10), seed=3)
    gs = []
    for dt in (torch.float64, torch.float32):
        o = OracleNet(rec, dt, quant='bf16')
        out, g = o.grads(x0, y, tau=0.7)
        gs.append([g[(p, r, k, id(t))].double().numpy() for (p, r, k, t) in o.trainable])
    worst = 0
    for a, b in zip(*gs):
        n = np.linalg.norm(a)
        if n > 1e-6: worst = max(worst, np.linalg.norm(a - b) / n)
    print(hy, 'quant-oracle fp64 vs fp32 accumulation: worst per-tensor rel diff %.3f' % worst)

hy = dict(k_cpt=4e-9, x0_shape=(16, 16, 1), n_cls=5)
net = randomize_routers(tiny_net('ac', seed=0, **hy))
rec = record_of(net)
x0, y = batch(24, x0_shape=(16, 16, 1), n_cls=5, seed=3)
gs = []
for dt in (torch.float64, torch.float32):
    o = OracleNet(rec, dt, quant='bf16')
    out, g = o.grads(x0, y, tau=0.7)
    gs.append([g[(p, r, k, id(t))].double().numpy() for (p, r, k, t) in o.trainable])
    names = [(p, r, k) for (p, r, k, t) in o.trainable]
sp = sorted(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30) for a, b in zip(*gs) if np.linalg.norm(a) > 1e-6)
print('per-tensor spreads: median %.3f, quartiles %.3f %.3f, max %.3f' % (np.median(sp), sp[len(sp) // 4], sp[3 * len(sp) // 4], sp[-1]))
for n_, a, b in list(zip(names, *gs))[:8]:
    print(n_, '%.3f' % (np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)))
