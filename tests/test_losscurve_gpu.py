"""bf16 (tcgen05) training against fp32 (the reference's arithmetic) over 300 steps of the full cifar10-ac
net on LEARNABLE synthetic data (class-dependent mean image + noise): the two loss curves must stay inside
a stated band and both must learn.  This is the end-to-end evidence behind running the drivers in bf16
(`--precision bf16`): per-step gradients differ by 10-20 % at random initialisation,
the trajectories do not diverge."""
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


def _curve(prec, steps=300, B=128):
    import arch_and_hypers as ah
    from lib import layer_types
    layer_types.seed(0)
    net = ah.ac_chain(k_cpt=4e-9)((32, 32, 3), (10,)).configure(precision=prec, graphs=True)
    rng = np.random.default_rng(0)
    means = rng.random((10, 32, 32, 3)).astype(np.float32)             # one mean image per class
    eng = net._get_engine()
    plan = eng._plan(B, True, True)
    losses = []
    for t in range(steps):
        cls = rng.integers(0, 10, B)
        x0 = np.clip(0.5 * means[cls] + 0.5 * rng.random((B, 32, 32, 3), dtype=np.float32), 0, 1).astype(np.float32)
        y = np.eye(10, dtype=np.float32)[cls]
        net.train.run({net.x0: x0, net.y: y, net.mode: 'tr', net.λ_lrn: ah.λ_lrn(t), net.τ: ah.τ_ds(t)})
        losses.append(float(plan.c_data.mean()))                       # per-example objective of the step (syncs)
    # accuracy of the trained net on fresh data from the same distribution
    cls = rng.integers(0, 10, 512)
    x0 = np.clip(0.5 * means[cls] + 0.5 * rng.random((512, 32, 32, 3), dtype=np.float32), 0, 1).astype(np.float32)
    st = net.eval_stats({net.x0: x0, net.y: np.eye(10, dtype=np.float32)[cls], net.τ: ah.τ_ds(steps)})
    return np.array(losses), float(st[(net, 'acc')].mean())


def test_bf16_and_fp32_loss_curves_stay_together():
    l32, acc32 = _curve('fp32')
    l16, acc16 = _curve('bf16')
    sm = lambda v: np.convolve(v, np.ones(20) / 20, mode='valid')      # 20-step moving average
    a, b = sm(l32), sm(l16)
    gap = np.abs(a - b) / a
    print('LOSSCURVE fp32 %.4f -> %.4f (acc %.3f) | bf16 %.4f -> %.4f (acc %.3f) | max smoothed gap %.3f, final gap %.3f'
          % (a[0], a[-1], acc32, b[0], b[-1], acc16, gap.max(), gap[-1]))
    assert np.isfinite(l32).all() and np.isfinite(l16).all()
    assert a[-1] < 0.5 * a[0] and b[-1] < 0.5 * b[0]                   # both learn
    assert gap.max() < 0.25 and gap[-1] < 0.15                         # stated band
    assert acc32 > 0.9 and acc16 > 0.9 and abs(acc32 - acc16) < 0.05
