"""The oracle against itself: PyTorch restatement vs independent NumPy fp64
forward vs finite differences.  (The reference ships no golden vectors --
SURVEY section 4 -- so this is the strongest pin available: parity unpinned.)"""
import numpy as np
import pytest
import torch

from oracle import np_ref
from oracle.torch_ref import OracleNet, first_argmax
from util import batch, node_paths, randomize_routers, record_of, tiny_net


@pytest.mark.parametrize('kind,hy', [
    ('sr', {}), ('ac', dict(k_cpt=4e-9)), ('cr', dict(k_cpt=4e-9)), ('cr', dict(k_cpt=1e-8, optimistic=True)),
    ('cr', dict(k_cpt=1e-8, use_cls_err=True)), ('actree', dict(k_cpt=2e-9)), ('ac', dict(dyn_k_cpt=True))])
@pytest.mark.parametrize('mode', ['tr', 'ev'])
def test_torch_vs_numpy_forward(kind, hy, mode):
    net = tiny_net(kind, **hy)
    if kind != 'sr':
        randomize_routers(net)
    rec = record_of(net)
    x0, y = batch(12)
    kc = np.random.default_rng(3).choice([0.0, 1e-9, 6.4e-8], 12) if hy.get('dyn_k_cpt') else None
    o = OracleNet(rec, torch.float64)
    a = o.forward(x0, y, mode, tau=0.7, k_cpt=kc)
    b = np_ref.forward(rec, x0, y, mode, tau=0.7, k_cpt=kc)
    assert abs(float(a.c_tot.detach()) - b['c_tot']) < 1e-10 * max(1, abs(b['c_tot']))
    for p in a.order:
        na, nb = a.nodes[p], b['nodes'][p]
        if kind != 'sr':
            np.testing.assert_allclose(na.p_tr.detach().numpy(), nb['p_tr'], rtol=1e-10, atol=1e-14)
            np.testing.assert_array_equal(na.p_ev.numpy(), nb['p_ev'])
        if not na.rec['sinks']:
            np.testing.assert_allclose(na.c_err.detach().numpy(), nb['c_err'], rtol=1e-10)
            np.testing.assert_array_equal(na.delta_cor.numpy(), nb['d_cor'])


def test_first_argmax_ties():
    x = torch.tensor([[0., 0., 0.], [1., 3., 3.], [2., 1., 2.]])
    assert first_argmax(x, 1).tolist() == [0, 1, 0]


@pytest.mark.parametrize('kind,hy', [('sr', {}), ('ac', dict(k_cpt=1e-7)), ('cr', dict(k_cpt=1e-7))])
def test_autograd_vs_finite_differences(kind, hy):
    """d c_tot / d theta from the torch oracle == central differences of the
    NumPy fp64 forward, for a handful of entries of every parameter family."""
    net = tiny_net(kind, **hy)
    if kind != 'sr':
        randomize_routers(net)
    rec = record_of(net)
    x0, y = batch(6)
    o = OracleNet(rec, torch.float64)
    out, g = o.grads(x0, y, tau=0.8)
    rng = np.random.default_rng(0)
    base = np_ref.forward(rec, x0, y, 'tr', tau=0.8)['nodes']      # values under tf.stop_gradient

    # walk records in the same order as OracleNet._collect to pair tensors with arrays
    pairs = []

    def grab(r):
        if r is None:
            return
        for k in r['params']:
            if k not in ('m_avg', 'v_avg'):
                pairs.append((r['params'], k))
        for c in r['comps']:
            grab(c)

    def collect(r):
        grab(r); grab(r['router'])
        for s in r['sinks']:
            collect(s)
    collect(rec['root'])
    assert len(pairs) == len(o.trainable)
    checked = 0
    for (pd, k), (path, role, key, t) in zip(pairs, o.trainable):
        assert k == key
        gr = g[(path, role, key, id(t))].numpy()
        if rng.random() > 0.25:
            continue
        idx = tuple(rng.integers(0, s) for s in pd[k].shape)
        orig = pd[k][idx]
        h = 1e-4 * max(1.0, abs(float(orig)))
        pd[k] = pd[k].astype(np.float64)
        vals = []
        for sgn in (+1, -1):
            pd[k][idx] = orig + sgn * h
            vals.append(np_ref.forward(rec, x0, y, 'tr', tau=0.8, frozen=base)['c_tot'])
        pd[k][idx] = orig
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - gr[idx]) <= 2e-4 * max(1e-3, abs(fd)) + 1e-9, (path, role, key, idx, fd, gr[idx])
        checked += 1
    assert checked >= 5


def test_train_step_momentum_and_talr():
    """TALR scale = 1/sqrt(mean p_tr^2) and a <- mu a + g; theta <- theta - lr a."""
    net = randomize_routers(tiny_net('ac', k_cpt=1e-8))
    rec = record_of(net)
    x0, y = batch(8)
    o = OracleNet(rec, torch.float64)
    before = [t.detach().clone() for _, _, _, t in o.trainable]
    out, g = o.grads(x0, y, tau=1.0)
    o2 = OracleNet(rec, torch.float64)
    res = o2.train_step(x0, y, lr=0.1, mu=0.9, tau=1.0)
    for (path, role, key, t), t2, b in zip(o.trainable, [t for *_, t in o2.trainable], before):
        s = 1 / np.sqrt(float((res.nodes[path].p_tr.detach() ** 2).mean()))
        if t2.grad is None:
            assert torch.equal(t2.detach(), b)
            continue
        exp = b - 0.1 * s * g[(path, role, key, id(t))]
        np.testing.assert_allclose(t2.detach().numpy(), exp.numpy(), rtol=1e-9, atol=1e-12)


def test_routing_invariants():
    """sum_children p_tr = parent p_tr; sum_leaves p_ev = 1; p_tr >= floor."""
    net = randomize_routers(tiny_net('actree', k_cpt=1e-9))
    rec = record_of(net)
    x0, y = batch(16)
    out = OracleNet(rec, torch.float64).forward(x0, y, 'tr', tau=0.3)
    nodes = out.nodes
    leaves = [p for p in out.order if not nodes[p].rec['sinks']]
    np.testing.assert_allclose(sum(nodes[p].p_ev for p in leaves).numpy(), 1.0)
    np.testing.assert_allclose(sum(nodes[p].p_tr for p in leaves).detach().numpy(), 1.0, rtol=1e-12)
    for p in out.order:
        kids = [(p + '/' if p else '') + str(i) for i in range(len(nodes[p].rec['sinks']))]
        if kids:
            tot = sum(nodes[k].p_tr for k in kids) if len(kids) > 1 else nodes[kids[0]].p_tr
            np.testing.assert_allclose(tot.detach().numpy(), nodes[p].p_tr.detach().numpy(), rtol=1e-12)
