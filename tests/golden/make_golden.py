#!/usr/bin/env python3
"""Generate the committed golden vectors (tests/golden/*.npz).

The reference ships no golden vectors and cannot run here (TensorFlow <= 0.12,
SURVEY F6), so these are produced by the fp64 oracle (oracle/torch_ref.py,
cross-checked against oracle/np_ref.py) on seeded nets and inputs:
PARITY UNPINNED with respect to the reference itself.  They pin (a) the oracle
against accidental drift (tests/test_golden.py, CPU) and (b) the CUDA path on
the GPU box, where /root/reference and this generator's environment are absent.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import np_ref  # noqa: E402
from oracle.torch_ref import OracleNet  # noqa: E402
from util import batch, randomize_routers, record_of, tiny_net  # noqa: E402

CASES = {
    'sr': ('sr', {}),
    'ac': ('ac', dict(k_cpt=4e-9)),
    'cr': ('cr', dict(k_cpt=4e-9)),
    'cr_opt': ('cr', dict(k_cpt=1e-8, optimistic=True)),
    'actree': ('actree', dict(k_cpt=2e-9)),
}
B, TAU = 24, 0.7


def build(name):
    kind, hy = CASES[name]
    net = tiny_net(kind, seed=0, **hy)
    if kind != 'sr':
        randomize_routers(net)
    return net


def golden(name):
    net = build(name)
    rec = record_of(net)
    x0, y = batch(B, seed=3)
    o = OracleNet(rec, torch.float64)
    out, g = o.grads(x0, y, tau=TAU)
    ref = np_ref.forward(rec, x0, y, 'tr', tau=TAU)
    assert abs(float(out.c_tot.detach()) - ref['c_tot']) < 1e-10
    d = {'c_tot': np.float64(out.c_tot.detach()), 'B': B, 'tau': TAU}
    leaves = [p for p in out.order if not out.nodes[p].rec['sinks']]
    d['leaf_paths'] = np.array(leaves)
    d['leaf_c_err'] = np.stack([out.nodes[p].c_err.detach().numpy() for p in leaves])
    d['leaf_logits'] = np.stack([out.nodes[p].comps[1].x.detach().numpy() for p in leaves])
    d['leaf_d_cor'] = np.stack([out.nodes[p].delta_cor.numpy() for p in leaves])
    if out.nodes[''].p_tr is not None:
        d['p_tr'] = np.stack([out.nodes[p].p_tr.detach().numpy() for p in out.order])
        d['p_ev'] = np.stack([out.nodes[p].p_ev.numpy() for p in out.order])
        sw = [p for p in out.order if len(out.nodes[p].rec['sinks']) > 1]
        d['switch_paths'] = np.array(sw)
        for i, p in enumerate(sw):
            d['router_logits_%d' % i] = out.nodes[p].router.x.detach().numpy()
        d['leaf_counts'] = np.array([out.nodes[p].p_ev.sum().item() for p in leaves])
    # gradients: norm of every trainable tensor (in the oracle's enumeration order) + a checksum
    d['grad_norms'] = np.array([g[(p, r, k, id(t))].norm().item() for p, r, k, t in o.trainable])
    w = np.random.default_rng(7)
    d['grad_proj'] = np.array([float((g[(p, r, k, id(t))] * torch.tensor(w.standard_normal(t.shape))).sum())
                               for p, r, k, t in o.trainable])
    # three training steps: parameter checksum
    o2 = OracleNet(record_of(build(name)), torch.float64)
    for t in range(3):
        xb, yb = batch(16, seed=10 + t)
        o2.train_step(xb, yb, lr=0.05 / 2 ** t, mu=0.9, tau=1.0 / 2 ** (t / 2))
    d['theta_norms_after_3_steps'] = np.array([t.detach().norm().item() for *_, t in o2.trainable])
    return d


if __name__ == '__main__':
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **golden(name))
        print('wrote', name)
