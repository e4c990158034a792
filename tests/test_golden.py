"""Golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU: the oracle reproduces them (pins NumPy's seeded streams, the parameter
initialiser and the oracle itself).  GPU: the CUDA path, through
net.train.run / the C ABI, matches them without needing /root/reference or the
oracle's intermediate tensors."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden as mg  # noqa: E402
from util import batch, node_paths, rel_err  # noqa: E402

NAMES = sorted(mg.CASES)


def load(name):
    return np.load(os.path.join(HERE, 'golden', name + '.npz'), allow_pickle=False)


@pytest.mark.parametrize('name', NAMES)
def test_oracle_reproduces_golden(name):
    g = load(name)
    d = mg.golden(name)
    for k in g.files:
        if g[k].dtype.kind in 'US':
            assert list(g[k]) == list(d[k])
        else:
            np.testing.assert_allclose(d[k], g[k], rtol=1e-9, atol=1e-12, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_path_matches_golden_fp32(name):
    g = load(name)
    net = mg.build(name).configure(precision='fp32')
    eng = net._get_engine()
    x0, y = batch(int(g['B']), seed=3)
    feed = {net.x0: x0, net.y: y}
    if net.dynamic:
        feed[net.τ] = float(g['tau'])
    eng.train_step(feed, update=False)
    torch.cuda.synchronize()
    plan = eng._plan(int(g['B']), True, True)
    paths = [p for p, _ in node_paths(net)]
    assert abs(eng.c_tot(plan) - float(g['c_tot'])) < 1e-3 * abs(float(g['c_tot']))
    for i, p in enumerate(g['leaf_paths']):
        r = plan.reg[paths.index(str(p))]
        assert rel_err(r.Z.cpu().numpy(), g['leaf_logits'][i]) < 1e-3
        assert rel_err(r.c_err.cpu().numpy(), g['leaf_c_err'][i]) < 1e-3
        np.testing.assert_array_equal(r.d_cor.cpu().numpy(), g['leaf_d_cor'][i])
    if 'p_ev' in g.files:
        p_ev = plan.p_ev.cpu().numpy()
        np.testing.assert_array_equal(p_ev, g['p_ev'])                      # routing decisions, bit-exact
        leaves = [paths.index(str(p)) for p in g['leaf_paths']]
        np.testing.assert_array_equal(p_ev[leaves].sum(1), g['leaf_counts'])   # per-node example counts
        assert rel_err(plan.p_tr.cpu().numpy(), g['p_tr']) < 1e-3
    grads = eng.grads_numpy(with_l2=True)
    norms = np.array([np.linalg.norm(grads[p]) for p in eng.tparams])
    w = np.random.default_rng(7)
    proj = np.array([float((grads[p] * w.standard_normal(p.shape)).sum()) for p in eng.tparams])
    big = g['grad_norms'] > 1e-6
    np.testing.assert_allclose(norms[big], g['grad_norms'][big], rtol=1e-3)
    np.testing.assert_allclose(proj[big], g['grad_proj'][big], rtol=2e-3, atol=1e-3 * np.abs(g['grad_proj']).max())
    # three optimiser steps
    net2 = mg.build(name).configure(precision='fp32')
    for t in range(3):
        xb, yb = batch(16, seed=10 + t)
        f = {net2.x0: xb, net2.y: yb, net2.mode: 'tr', net2.λ_lrn: 0.05 / 2 ** t}
        if net2.dynamic:
            f[net2.τ] = 1.0 / 2 ** (t / 2)
        net2.train.run(f)
    e2 = net2._get_engine()
    th = np.array([float(e2._buf(p).double().norm()) for p in e2.tparams])
    ref = g['theta_norms_after_3_steps']
    np.testing.assert_allclose(th, ref, rtol=2e-3, atol=2e-4)
