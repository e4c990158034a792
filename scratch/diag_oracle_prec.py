import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200')]
import numpy as np, torch
from lib import layer_types, serdes
import arch_and_hypers as ah
from oracle.torch_ref import OracleNet
torch.set_num_threads(8)
for B in (32, 48):
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
    gs = []
    for dt in (torch.float64, torch.float32):
        o = OracleNet(copy.deepcopy(rec), dt)
        out, g = o.grads(x0, y, tau=0.8)
        gs.append([g[(p, r, k, id(t))].double().numpy() for (p, r, k, t) in o.trainable])
        names = [(p, k) for (p, r, k, t) in o.trainable]
    tot = sum((a ** 2).sum() for a in gs[0]) ** 0.5
    errs = sorted(((float(np.sqrt(((a - b) ** 2).sum())) / tot, n) for a, b, n in zip(gs[0], gs[1], names)), reverse=True)
    print('B', B, 'fp32-vs-fp64 oracle: top errors', errs[:3])
