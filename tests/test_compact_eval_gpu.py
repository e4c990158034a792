"""Compacted 'ev'-mode evaluation (lib/compact_eval.py): the per-switch compaction and leaf
statistics kernels against NumPy, and the evaluator against (i) the dense engine path -- bit-exact on every
p_ev-weighted statistic, because each example's logits do not depend on which other examples share its batch
-- and (ii) the oracle's `state_tensors` (scripts/train-nets:111-130, scripts/lib/desc.py:10-22)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle.torch_ref import OracleNet  # noqa: E402
from util import batch, node_paths, randomize_routers, record_of, tiny_net  # noqa: E402

pytestmark = pytest.mark.gpu


def L():
    from lib import _cabi
    return _cabi.lib()


def vp(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize('n,ns', [(1, 2), (37, 2), (1024, 3), (5000, 2), (2049, 8)])
def test_route_compact_matches_numpy(n, ns):
    rng = np.random.default_rng(n)
    R = rng.standard_normal((n, ns)).astype(np.float32)
    R[rng.random(n) < 0.2] = 0.0                                    # ties resolve to the first maximum
    parent = rng.permutation(3 * n)[:n].astype(np.int32)
    cap = n + 5
    Rd, pd = torch.from_numpy(R).cuda(), torch.from_numpy(parent).cuda()
    dec = torch.zeros(n, dtype=torch.int32, device='cuda')
    pos = torch.full((ns, cap), -1, dtype=torch.int32, device='cuda'); orig = torch.full_like(pos, -1)
    cnt = torch.zeros(ns, dtype=torch.int32, device='cuda')
    L().route_compact(vp(Rd), ns, ns, n, vp(pd), cap, vp(dec), vp(pos), vp(orig), vp(cnt), None)
    torch.cuda.synchronize()
    want = R.argmax(1)
    np.testing.assert_array_equal(dec.cpu().numpy(), want)
    for s in range(ns):
        rows = np.nonzero(want == s)[0]
        assert int(cnt[s]) == len(rows)
        np.testing.assert_array_equal(pos[s, :len(rows)].cpu().numpy(), rows)            # ascending, order preserving
        np.testing.assert_array_equal(orig[s, :len(rows)].cpu().numpy(), parent[rows])


def test_leaf_stats_matches_numpy():
    rng = np.random.default_rng(3)
    n_par, n_cls, B = 700, 10, 2000
    Z = rng.standard_normal((n_par, 16)).astype(np.float32)
    y = np.eye(n_cls, dtype=np.float32)[rng.integers(0, n_cls, B)]
    pos = np.sort(rng.choice(n_par, 300, replace=False)).astype(np.int32)
    orig = rng.choice(B, 300, replace=False).astype(np.int32)
    out = torch.zeros(2 + 2 * n_cls, dtype=torch.float64, device='cuda')
    cnt = torch.tensor([300], dtype=torch.int32, device='cuda')
    args = [torch.from_numpy(a).cuda() for a in (Z, y, pos, orig)]
    for rep in range(2):                                             # accumulates
        L().leaf_stats(vp(args[0]), 16, n_cls, vp(args[1]), vp(args[2]), vp(args[3]), vp(cnt), 300, vp(out), None)
    torch.cuda.synchronize()
    cor = Z[pos, :n_cls].argmax(1) == y[orig].argmax(1)
    want = np.concatenate([[cor.sum(), (~cor).sum()], (cor[:, None] * y[orig]).sum(0), ((~cor)[:, None] * y[orig]).sum(0)])
    np.testing.assert_array_equal(out.cpu().numpy(), 2 * want)


def _dense_means(net, x0, y, tau, kc=None):
    feed = {net.x0: x0, net.y: y, net.τ: tau}
    if kc is not None:
        feed[net.k_cpt] = kc
    st = net.eval_stats(feed)
    return {k: np.asarray(v, np.float64).mean(0) for k, v in st.items()}


EXACT = ('p_cor', 'p_inc', 'p_cor_by_cls', 'p_inc_by_cls')


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('kind,hy', [('ac', dict(k_cpt=4e-9)), ('cr', dict(k_cpt=4e-9)), ('actree', dict(k_cpt=2e-9)),
                                     ('ac', dict(dyn_k_cpt=True))])
def test_compact_evaluator_equals_the_dense_path(kind, hy, prec):
    net = randomize_routers(tiny_net(kind, seed=2, **hy)).configure(precision=prec)
    x0, y = batch(200, seed=9)
    for t in range(2):                                   # non-trivial running BN moments
        f = {net.x0: x0[:64], net.y: y[:64], net.τ: 0.5, net.mode: 'tr', net.λ_lrn: 0.01}
        if hy.get('dyn_k_cpt'):
            f[net.k_cpt] = np.full(64, 4e-9, np.float32)
        net.train.run(f)
    kc = [1.6e-8] if hy.get('dyn_k_cpt') else None
    dense = _dense_means(net, x0, y, 0.5, kc)
    ev = net.compact_evaluator(128)                      # two batches: 128 + 72
    ev.reset()
    for i in range(0, 200, 128):
        ev.run_batch(x0[i:i + 128], y[i:i + 128], k_cpt=kc)
    got = ev.result()
    assert got[(net, 'acc')] == pytest.approx(dense[(net, 'acc')], abs=1e-12)
    assert got[(net, 'moc')] == pytest.approx(dense[(net, 'moc')], rel=1e-12)
    for leaf in net.leaves:
        for name in EXACT:
            np.testing.assert_allclose(got[(leaf, name)], dense[(leaf, name)], rtol=0, atol=1e-12)
    assert not any(name in ('c_err', 'p_tr', 'x_rte') for _, name in got)
    # every example is counted at exactly one leaf, and fewer node visits than the dense pass
    assert sum(got[(leaf, 'p_cor')] + got[(leaf, 'p_inc')] for leaf in net.leaves) == pytest.approx(1.0, abs=1e-12)
    assert ev.visits.sum() < 200 * len(ev.visits)


def test_compact_evaluator_matches_the_oracle_on_the_full_net():
    """ac_chain at the real architecture, 512 examples in batches of 256, against OracleNet.state"""
    import arch_and_hypers as ah
    from lib import layer_types
    layer_types.seed(0)
    net = ah.ac_chain(k_cpt=4e-9)((32, 32, 3), (10,)).configure(precision='fp32')
    rng = np.random.default_rng(1)
    for l in net.layers:
        if l.router is not None:
            w = l.router.comps[-1].params.w
            w.assign((0.5 * rng.standard_normal(w.shape)).astype(np.float32))
    x0 = rng.random((512, 32, 32, 3)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 512)]
    for t in range(12):                                  # move the running BN moments away from (0, 1)
        net.train.run({net.x0: x0[:128], net.y: y[:128], net.τ: 1.0, net.mode: 'tr', net.λ_lrn: 0.0})
    o = OracleNet(record_of(net), torch.float64)
    ref, out = o.state(x0, y, tau=0.5)
    ev = net.compact_evaluator(256)
    ev.reset()
    for i in range(0, 512, 256):
        ev.run_batch(x0[i:i + 256], y[i:i + 256])
    got = ev.result()
    paths = node_paths(net)
    # an example whose decision margin is inside fp32 rounding somewhere may land at another leaf: allow 1 %
    assert abs(got[(net, 'acc')] - ref[('net', 'acc')].mean()) <= 0.01
    assert abs(got[(net, 'moc')] - ref[('net', 'moc')].mean()) <= 0.01 * ref[('net', 'moc')].mean()
    n_diff = 0.0
    for path, l in paths:
        if (path, 'p_cor') in ref:
            for name in EXACT:
                n_diff += np.abs(np.asarray(got[(l, name)]) - ref[(path, name)].mean(0)).sum()
    print('compact vs oracle: total |difference| of the leaf statistics %.4f, node visits %d of %d dense; per node %s'
          % (n_diff, ev.visits.sum(), 512 * len(ev.visits), ev.visits.tolist()))
    assert n_diff <= 0.02
