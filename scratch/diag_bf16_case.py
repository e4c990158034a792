import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from oracle.torch_ref import OracleNet
from util import batch, randomize_routers, record_of, tiny_net
ncls = int(os.environ.get('NCLS', 5))
ch = int(os.environ.get('CH', 3))
hy = dict(k_cpt=4e-9, n_cls=ncls, x0_shape=(16, 16, ch))
net = randomize_routers(tiny_net('ac', seed=0, **hy)).configure(precision='bf16')
rec = record_of(net)
x0, y = batch(24, x0_shape=(16, 16, ch), n_cls=ncls, seed=3)
o = OracleNet(rec, torch.float64, quant='bf16')
out, g_ref = o.grads(x0, y, tau=0.7)
eng = net._get_engine()
eng.train_step({net.x0: x0, net.y: y, net.τ: 0.7}, update=False)
torch.cuda.synchronize()
g = eng.grads_numpy(with_l2=True)
plan = eng._plan(24, True, True)
for p, (path, role, key, t) in zip(eng.tparams, o.trainable):
    ref = g_ref[(path, role, key, id(t))].numpy()
    n = np.linalg.norm(ref)
    print('%-8s %-7s %-9s |ref| %.3e  rel err %.3f' % (path, role, key, n, np.linalg.norm(g[p] - ref) / max(n, 1e-30)))
