import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200')]
import numpy as np, torch
from lib import layer_types, serdes
import arch_and_hypers as ah
from oracle.torch_ref import OracleNet
B = int(os.environ.get('B', 64))
layer_types.seed(0)
net = ah.ac_chain(k_cpt=4e-9)((32, 32, 3), (10,))
rng = np.random.default_rng(0)
for l in net.layers:
    if l.router is not None:
        w = l.router.comps[-1].params.w
        w.assign((0.5 * rng.standard_normal(w.shape)).astype(np.float32))
rec = copy.deepcopy(serdes.encode_net(net))
x0 = rng.random((B, 32, 32, 3)).astype(np.float32)
y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
layer_types.seed(0)
n2 = serdes.decode_net(copy.deepcopy(rec)).configure(precision='fp32')
eng = n2._get_engine()
eng.train_step({n2.x0: x0, n2.y: y, n2.τ: 0.8}, update=False)
torch.cuda.synchronize()
o = OracleNet(copy.deepcopy(rec), torch.float64)
out, g = o.grads(x0, y, tau=0.8)
got = eng.grads_numpy(with_l2=True)
tot = sum(float((g[(p_, r, k, id(t))].numpy() ** 2).sum()) for (p_, r, k, t) in o.trainable) ** 0.5
print('global grad norm', tot)
rows = []
for p, (path, role, key, t) in zip(eng.tparams, o.trainable):
    ref = g[(path, role, key, id(t))].numpy()
    err = float(np.sqrt(((got[p] - ref) ** 2).sum()))
    rows.append((err / tot, path, role, key, float(np.linalg.norm(ref)), float(np.linalg.norm(got[p]))))
for r in sorted(rows, reverse=True)[:12]:
    print('%.2e  %-10s %-7s %-10s |ref| %.3e |got| %.3e' % r)
