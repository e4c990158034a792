"""The drivers on hardware: `train-nets` and `train-adaptive-nets` run as subprocesses on synthetic data
(datasets cannot be downloaded here) and the files they write are checked against what the reference's
consumers read (scripts/make-nlds:46-63, make-routing-hists:19-27, make-acc-eff-plots:25-28; schema:
scripts/lib/desc.py:24-36, scripts/train-nets:144-157, scripts/train-adaptive-nets:102-106)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'multipath-nn_b200')
pytestmark = pytest.mark.gpu


def _run(script, args, cwd):
    r = subprocess.run([sys.executable, os.path.join(PKG, script)] + args, cwd=cwd, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def _check_stats(stats, dynamic, n_cls=10):
    """the access patterns of the plotting scripts"""
    assert set(stats) >= {'type', 'stats_tr', 'stats_ts', 'root'}
    for split in ('stats_tr', 'stats_ts'):
        assert 0.0 <= stats[split]['acc'] <= 1.0 and stats[split]['moc'] > 0        # make-acc-eff-plots:25-28
    node = stats['root']['sinks'][0]                    # first ReConvMax (make-nlds:46-49)
    assert stats['root']['name'] == 'ToPyramid' and node['name'] == 'ReConvMax'
    n_leaves = 0
    p_total = 0.0
    while True:
        leaf = node['sinks'][0]                          # sinks[0] = the stage's classifier (make-nlds:59-63)
        assert leaf['name'] == 'LogReg'
        st = leaf['stats_ts']
        for k in ('p_cor', 'p_inc', 'c_err'):
            assert isinstance(st[k], float), (k, type(st[k]))
        for k in ('p_cor_by_cls', 'p_inc_by_cls'):
            assert isinstance(st[k], list) and len(st[k]) == n_cls
        if dynamic:
            assert isinstance(st['p_tr'], float)
        p_total += st['p_cor'] + st['p_inc']
        n_leaves += 1
        if len(node['sinks']) < 2:
            break
        assert 'x_rte' in node['stats_ts']               # switch statistic (make-routing-hists)
        node = node['sinks'][1]                          # sinks[1:] = continuations
        assert node['name'] == 'ReConvMax'
    assert abs(p_total - 1.0) < 1e-6                     # every test example is counted at exactly one leaf
    return n_leaves


def test_train_nets_cifar10_ac_writes_the_reference_layout(tmp_path):
    out = _run('train-nets', ['cifar10-ac', '--synthetic', '--n-iter', '20', '--t-log', '10', '--nets', '0',
                              '--t-ckpt', '10'], str(tmp_path))
    d = tmp_path / 'nets' / 'cifar10-ac'
    for f in ('0000.npy', '0000-stats.npy', '0000-log.txt', '0000-stats/00000010.npy', '0000-stats/00000020.npy',
              '0000-ckpt.npy'):
        assert (d / f).exists(), f
    stats = np.load(d / '0000-stats.npy', allow_pickle=True)[()]
    assert stats['type'] == 'ActorNet'
    assert _check_stats(stats, dynamic=True) == 8
    assert 'ReConvMax' in out and 'LogReg' in out                      # the rendered tree is printed (train-nets:155)
    net = np.load(d / '0000.npy', allow_pickle=True)[()]
    assert net['type'] == 'ActorNet' and net['root']['name'] == 'ToPyramid'
    # resume continues from the checkpoint instead of starting over
    out = _run('train-nets', ['cifar10-ac', '--synthetic', '--n-iter', '22', '--t-log', '11', '--nets', '0',
                              '--t-ckpt', '10', '--resume'], str(tmp_path))
    assert 'resumed' in out and 'step 20' in out


def test_train_nets_sr_and_cr(tmp_path):
    _run('train-nets', ['mnist-sr', '--synthetic', '--n-iter', '6', '--t-log', '6', '--nets', '7',
                        '--precision', 'fp32'], str(tmp_path))
    stats = np.load(tmp_path / 'nets' / 'mnist-sr' / '0007-stats.npy', allow_pickle=True)[()]
    assert stats['type'] == 'SRNet' and 0.0 <= stats['stats_ts']['acc'] <= 1.0
    _run('train-nets', ['cifar10-cr', '--synthetic', '--n-iter', '6', '--t-log', '6', '--nets', '3'], str(tmp_path))
    stats = np.load(tmp_path / 'nets' / 'cifar10-cr' / '0003-stats.npy', allow_pickle=True)[()]
    assert stats['type'] == 'CriticNet' and _check_stats(stats, dynamic=True) == 8


def test_train_adaptive_nets_writes_one_stats_file_per_k_cpt(tmp_path):
    """length-1 k_cpt feed `[k]` at statistics time (train-adaptive-nets:102-105) through mean_net_state"""
    _run('train-adaptive-nets', ['hybrid-ac-dynkcpt', '--synthetic', '--n-iter', '20'], str(tmp_path))
    d = tmp_path / 'nets' / 'hybrid-ac-dynkcpt'
    assert (d / 'net.npy').exists()
    mocs = []
    for i in range(8):
        stats = np.load(d / ('%.4i-stats.npy' % i), allow_pickle=True)[()]
        assert _check_stats(stats, dynamic=True) == 8
        mocs.append(stats['stats_ts']['moc'])
    assert all(np.isfinite(mocs))
