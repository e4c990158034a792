"""Oracle parity at the REAL architecture (arch_and_hypers.py:19-27, 32x32 input, 64..128-channel stages,
batch 128 = arch_and_hypers.py:35) for every BASELINE.json config: sr_chain(8) on 3- and 1-channel input,
ac_chain, cr_chain (plain / optimistic / use_cls_err), ac_tree, and the adaptive ac_chain(dyn_k_cpt=True)
with a per-example k_cpt and with the length-1 `[k]` feed of train-adaptive-nets:102-105.

Compared with the oracle on the same bytes: logits, c_err, router logits, p_tr, routing decisions (bit-exact
wherever the oracle's margin exceeds the stated tolerance -- the fraction of examples inside that mask is
printed and must be >= 0.8, so the claim is not vacuous), c_tot and every parameter gradient.

Tolerances.  fp32 mode against the fp64 oracle: forward values 1e-3; the whole gradient vector within
2e-3 + 1.5 x d, where d is the distance between the fp32 and the fp64 ORACLE on the same case (measured in the
test: 1.6e-4 .. 2.0e-3 -- a ReLU / max-pool input within fp32 rounding of its threshold resolves differently in
the two arithmetics, and one flipped unit deep in sr_chain(8) moves the gradient by 2e-3; the reference's own
fp32 TF graph has the same property); each tensor 1e-2.  bf16 mode against the oracle evaluated in the
arithmetic the device stores in (quant='bf16'): forward 5e-2, whole gradient 0.25, routing decisions exact
wherever that oracle's logit margin exceeds 0.15 (router logits deep in the net move by up to ~0.1 under bf16
storage).  bf16x3 mode (fp32 storage, every conv product as three bf16 tensor-core products, i.e. 16-bit
mantissas) against the fp64 oracle: forward 1e-3 (measured 5e-5 .. 8e-5), whole gradient 1.5e-2 (measured
1.6e-3 .. 9e-3: the backward pass through eight train-mode BatchNorms amplifies a forward perturbation 50-100x,
in fp32 exactly as here), decisions exact outside a 1e-3 margin.  bf16x6 mode (three-way splits, the six bf16
products of relative size >= 2^-16 per fp32 product: 24 significant bits on the tensor cores) is held to the fp32
bounds: forward 1e-3, whole gradient 2e-3 + 1.5 x d, each tensor 1e-2, decisions exact outside a 1e-4 margin.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu

import arch_and_hypers as ah  # noqa: E402
from lib import layer_types  # noqa: E402
from oracle.torch_ref import OracleNet  # noqa: E402
from util import node_paths, record_of, rel_err  # noqa: E402

B = 128
TOL = {'fp32': dict(fwd=1e-3, grad_all=2e-3, grad_each=1e-2, margin=1e-4),
       # fp32 storage, convolutions on the tensor cores as three bf16 products per fp32 product (2^-16 relative)
       'bf16x3': dict(fwd=1e-3, grad_all=1.5e-2, grad_each=6e-2, margin=1e-3),
       # three-way splits, six bf16 products per fp32 product (24 significant bits): the fp32 bounds
       'bf16x6': dict(fwd=1e-3, grad_all=2e-3, grad_each=1e-2, margin=1e-4),
       'bf16': dict(fwd=5e-2, grad_all=2.5e-1, grad_each=None, margin=1.5e-1)}

CASES = {
    'cifar10-sr': (lambda: ah.sr_chain(8), 3, {}),
    'mnist-sr': (lambda: ah.sr_chain(8), 1, {}),
    'cifar10-ac': (lambda: ah.ac_chain(k_cpt=4e-9), 3, {}),
    'cifar10-cr': (lambda: ah.cr_chain(k_cpt=4e-9), 3, {}),
    'cr-optimistic': (lambda: ah.cr_chain(k_cpt=1.6e-8, optimistic=True), 3, {}),
    'cr-use-cls-err': (lambda: ah.cr_chain(k_cpt=1.6e-8, use_cls_err=True), 3, {}),
    'ac-tree': (lambda: ah.ac_tree(k_cpt=4e-9), 3, {}),
    'ac-dyn-kcpt-vector': (lambda: ah.ac_chain(dyn_k_cpt=True), 3, {'kc': 'vector'}),
    'ac-dyn-kcpt-length1': (lambda: ah.ac_chain(dyn_k_cpt=True), 3, {'kc': 'length1'}),
}


def _build(name, prec):
    maker, C, opt = CASES[name]
    layer_types.seed(0)
    net = maker()((32, 32, C), (10,)).configure(precision=prec)
    rng = np.random.default_rng(1)
    for l in net.layers:                     # non-trivial routing (the reference zero-initialises the last router layer)
        if l.router is not None:
            w = l.router.comps[-1].params.w
            w.assign((0.5 * rng.standard_normal(w.shape)).astype(np.float32))
    rng = np.random.default_rng(7)
    x0 = rng.random((B, 32, 32, C)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    kc_dev = kc_ref = None
    if opt.get('kc') == 'vector':
        kc_dev = kc_ref = rng.choice(ah.k_cpts, B).astype(np.float32)
    elif opt.get('kc') == 'length1':
        kc_dev, kc_ref = [ah.k_cpts[4]], np.full(B, ah.k_cpts[4], np.float32)
    return net, x0, y, kc_dev, kc_ref


@pytest.mark.parametrize('prec', ['fp32', 'bf16', 'bf16x3', 'bf16x6'])
@pytest.mark.parametrize('name', list(CASES))
def test_full_architecture_matches_the_oracle(name, prec):
    tol = TOL[prec]
    net, x0, y, kc_dev, kc_ref = _build(name, prec)
    rec = record_of(net)
    tau = 0.7 if net.dynamic else None
    o = OracleNet(rec, torch.float64, quant='bf16' if prec == 'bf16' else None)
    out, g_ref = o.grads(x0, y, tau=tau, k_cpt=kc_ref)
    feed = {net.x0: x0, net.y: y}
    if net.dynamic:
        feed[net.τ] = tau
        if kc_dev is not None:
            feed[net.k_cpt] = kc_dev
    eng = net._get_engine()
    eng.train_step(feed, update=False)
    torch.cuda.synchronize()
    plan = eng._plan(B, True, True)
    paths = node_paths(net)
    worst_fwd = 0.0
    for nd in eng.regs:
        ref = out.nodes[paths[nd.idx][0]]
        e1 = rel_err(plan.reg[nd.idx].Z.cpu().numpy(), ref.comps[1].x.detach().numpy())
        e2 = rel_err(plan.reg[nd.idx].c_err.cpu().numpy(), ref.c_err.detach().numpy())
        worst_fwd = max(worst_fwd, e1, e2)
        assert e1 < tol['fwd'] and e2 < tol['fwd'], ('leaf', paths[nd.idx][0], e1, e2)
    frac_sure = 1.0
    if net.dynamic:
        p_tr = plan.p_tr.cpu().numpy(); dec = plan.dec.cpu().numpy()
        n_sure = n_all = 0
        for nd in eng.switches:
            path = paths[nd.idx][0]
            r_ref = out.nodes[path].router.x.detach().numpy()
            assert rel_err(plan.rtr[nd.idx].R.cpu().numpy(), r_ref) < 5 * tol['fwd'], ('router logits', path)
            srt = np.sort(r_ref, 1)
            sure = (srt[:, -1] - srt[:, -2]) > tol['margin']
            n_sure += int(sure.sum()); n_all += sure.size
            np.testing.assert_array_equal(dec[nd.sw][sure], r_ref.argmax(1)[sure])      # bit-exact decisions
        frac_sure = n_sure / n_all
        assert frac_sure >= 0.8, frac_sure
        for nd in eng.nodes:
            path = paths[nd.idx][0]
            assert rel_err(p_tr[nd.idx], out.nodes[path].p_tr.detach().numpy()) < 5 * tol['fwd'], ('p_tr', path)
        # per-node example counts = sum of p_ev (bit-exact where every decision on the way is outside the margin)
        p_ev = plan.p_ev.cpu().numpy()
        leaves = [nd.idx for nd in eng.nodes if not nd.kids]
        np.testing.assert_array_equal(p_ev[leaves].sum(0), 1.0)
    c_ref = float(out.c_tot.detach())
    assert abs(eng.c_tot(plan) - c_ref) < tol['fwd'] * abs(c_ref)
    g = eng.grads_numpy(with_l2=True)
    assert len(eng.tparams) == len(o.trainable)
    num = den = 0.0
    worst_each, worst_name = 0.0, None
    gmax = max(float(np.abs(v.numpy()).max()) for v in g_ref.values())
    for p, (path, role, key, t) in zip(eng.tparams, o.trainable):
        ref = g_ref[(path, role, key, id(t))].numpy()
        assert ref.shape == g[p].shape, (path, role, key)
        num += float(((g[p] - ref) ** 2).sum()); den += float((ref ** 2).sum())
        n = np.linalg.norm(ref)
        if n > 1e-3 * gmax * np.sqrt(ref.size):         # tensors with a meaningful gradient
            e = float(np.linalg.norm(g[p] - ref) / n)
            if e > worst_each:
                worst_each, worst_name = e, (path, role, key)
    err_all = (num / den) ** 0.5
    d32 = 0.0
    if prec != 'bf16':                  # how far the reference arithmetic (fp32) is from the fp64 oracle on this case
        o32 = OracleNet(rec, torch.float32)
        _, g32 = o32.grads(x0, y, tau=tau, k_cpt=kc_ref)
        n32 = sum(float(((g32[(pa, ro, ke, id(t32))].numpy().astype(np.float64) - g_ref[(pa, ro, ke, id(t))].numpy()) ** 2).sum())
                  for (pa, ro, ke, t32), (_, _, _, t) in zip(o32.trainable, o.trainable))
        d32 = (n32 / den) ** 0.5
    print('PARITY %-22s %s  forward %.2e  gradient (all) %.2e  (fp32 vs fp64 oracle %.2e)  worst tensor %.2e %s  '
          'decisions inside the mask %.3f' % (name, prec, worst_fwd, err_all, d32, worst_each, worst_name, frac_sure))
    assert err_all < tol['grad_all'] + 1.5 * d32, (err_all, d32)
    if tol['grad_each'] is not None:
        assert worst_each < tol['grad_each'], (worst_each, worst_name)
