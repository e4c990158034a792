"""Full-size nets (BASELINE.json configs 2-5 at the reference's batch 128): size-independent properties
of the CUDA path where the oracle is too slow to be the checker."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu

import arch_and_hypers as ah  # noqa: E402
from lib import layer_types  # noqa: E402

B = 128


def _data(n_cls=10, seed=0, C=3):
    rng = np.random.default_rng(seed)
    return rng.random((B, 32, 32, C)).astype(np.float32), np.eye(n_cls, dtype=np.float32)[rng.integers(0, n_cls, B)]


def _build(maker, prec, C=3, **kw):
    layer_types.seed(0)
    net = maker(**kw)((32, 32, C), (10,)).configure(precision=prec)
    rng = np.random.default_rng(1)
    for l in net.layers:                     # non-trivial routing (the reference zero-initialises the last router layer)
        if l.router is not None:
            w = l.router.comps[-1].params.w
            w.assign((0.5 * rng.standard_normal(w.shape)).astype(np.float32))
    return net


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('maker,kw', [(ah.ac_chain, dict(k_cpt=4e-9)), (ah.cr_chain, dict(k_cpt=4e-9)),
                                      (ah.ac_tree, dict(k_cpt=4e-9)), (ah.ac_chain, dict(dyn_k_cpt=True))])
def test_routing_invariants(maker, kw, prec):
    """every example reaches exactly one leaf; soft routing probabilities of the leaves sum to one; a
    node's p_ev is the sum of its children's; the statistics account for every example once"""
    net = _build(maker, prec, **kw)
    x0, y = _data()
    feed = {net.x0: x0, net.y: y, net.τ: 0.5, net.mode: 'ev'}
    if kw.get('dyn_k_cpt'):
        feed[net.k_cpt] = np.random.default_rng(2).choice(ah.k_cpts, B).astype(np.float32)
    eng = net._get_engine()
    res = eng.debug_forward(feed, mode='ev')
    leaves = [nd.idx for nd in eng.nodes if not nd.kids]
    np.testing.assert_array_equal(res.p_ev[leaves].sum(0), np.ones(B, np.float32))
    np.testing.assert_allclose(res.p_tr[leaves].sum(0), 1.0, rtol=0, atol=1e-5)
    for nd in eng.nodes:
        if nd.kids:
            np.testing.assert_array_equal(res.p_ev[list(nd.kids)].sum(0), res.p_ev[nd.idx])
    st = net.eval_stats(feed)
    total = sum(st[(nd.layer, 'p_cor')] + st[(nd.layer, 'p_inc')] for nd in eng.regs)
    np.testing.assert_array_equal(total, np.ones(B))
    assert 0.0 <= st[(net, 'acc')].mean() <= 1.0 and st[(net, 'moc')].min() > 0


@pytest.mark.parametrize('maker,kw,C', [(ah.sr_chain, dict(n_tf=8), 3), (ah.sr_chain, dict(n_tf=8), 1),
                                        (ah.ac_chain, dict(k_cpt=4e-9), 3), (ah.cr_chain, dict(k_cpt=4e-9), 3)])
def test_training_on_a_fixed_batch_reduces_the_objective(maker, kw, C):
    """40 steps of net.train.run on one batch (bf16, CUDA graphs): the objective falls and stays finite"""
    if 'n_tf' in kw:
        layer_types.seed(0)
        net = maker(kw['n_tf'])((32, 32, C), (10,)).configure(precision='bf16')
    else:
        net = _build(maker, 'bf16', C=C, **kw)
    x0, y = _data(C=C)
    eng = net._get_engine()
    feed = {net.x0: x0, net.y: y, net.mode: 'tr', net.λ_lrn: 0.02}
    if net.dynamic:
        feed[net.τ] = 1.0
    vals = []
    for t in range(40):
        net.train.run(feed)
        if t in (0, 39):
            torch.cuda.synchronize()
            vals.append(eng.c_tot(eng._plan(B, True, True)))
    assert np.isfinite(vals).all() and vals[1] < 0.9 * vals[0], vals


def test_bf16_forward_tracks_fp32_on_the_full_net():
    """logits of every leaf and the routing decisions of confident examples agree between the two modes"""
    x0, y = _data()
    outs = []
    for prec in ('fp32', 'bf16'):
        net = _build(ah.ac_chain, prec, k_cpt=4e-9)
        outs.append(net._get_engine().debug_forward({net.x0: x0, net.y: y, net.τ: 0.5}, mode='tr'))
    a, b = outs
    for i in a.logits:
        err = np.abs(a.logits[i] - b.logits[i]).max() / max(np.abs(a.logits[i]).max(), 1e-6)
        assert err < 5e-2, (i, err)
    for i in a.R:
        srt = np.sort(a.R[i], 1)
        sure = (srt[:, -1] - srt[:, -2]) > 0.25         # bf16 storage moves deep router logits by up to ~0.1
        np.testing.assert_array_equal(a.R[i].argmax(1)[sure], b.R[i].argmax(1)[sure])
