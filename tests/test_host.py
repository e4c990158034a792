"""CPU-only tests of the host side: the reference-facing API surface, the file
formats, the planner (dry run) and the C ABI's symbol table.  No GPU compute."""
import ctypes
import os

import numpy as np
import pytest
import torch

import arch_and_hypers as ah
from lib import _cabi, desc, layer_types, net_types, serdes
from lib.data import Dataset, synthetic_archive
from lib.net_types import params_list_rec
from util import record_of, tiny_net


def _n_params(net):
    return sum(p.value.size for l in net.layers
               for p in list(params_list_rec(l)) + list(params_list_rec(l.router)) if p.trainable)


def test_architecture_counts_match_survey():
    """680,798 / 531,594 trainable parameters, 20,699,872 / 20,551,424 MAC per image."""
    ac = ah.ac_chain(k_cpt=1e-9)((32, 32, 3), (10,))
    sr = ah.sr_chain(8)((32, 32, 3), (10,))
    assert _n_params(ac) == 680798 and _n_params(sr) == 531594
    assert sum(l.n_ops + (l.router.n_ops if l.router else 0) for l in ac.layers) == 20699872
    assert sum(l.n_ops for l in sr.layers) == 20551424
    assert len(list(ac.layers)) == 17 and len(list(ac.leaves)) == 8 and len(list(ac.switches)) == 7
    # router last layer is zero-initialised (sigma_w = 0): all decisions tie to sink 0 at init
    for l in ac.switches:
        assert not l.router.comps[-1].params.w.value.any()


def test_tree_constructors_work():
    """the reference's *-tree constructors raise NameError; here they build"""
    net = ah.ac_tree(k_cpt=1e-9)((32, 32, 3), (10,))
    assert len(list(net.leaves)) > 8 and max(len(l.sinks) for l in net.layers) == 3


def test_hyper_keys_follow_python_identifier_normalisation():
    """`ϵ` (U+03F5) in the reference's source becomes the attribute / record key U+03B5"""
    net = tiny_net('ac', k_cpt=1e-9)
    assert 'ε' in vars(net.hypers) and 'ϵ' not in vars(net.hypers)
    rec = serdes.encode_net(net)
    assert set(rec) == {'type', 'root', 'hypers', 'params'}
    assert {'σ_w', 'k_l2', 'n_chan', 'res'} <= set(rec['root']['sinks'][0]['router']['comps'][1]['hypers'])


@pytest.mark.parametrize('kind', ['sr', 'ac', 'cr', 'actree'])
def test_serdes_roundtrip(kind, tmp_path):
    net = tiny_net(kind, **({} if kind == 'sr' else dict(k_cpt=2e-9)))
    path = str(tmp_path / 'net.npy')
    serdes.write_net(path, net)
    raw = np.load(path, allow_pickle=True)
    assert raw.shape == () and raw.dtype == object          # np.save of a dict: 0-d object array
    net2 = serdes.read_net(path)

    def same(a, b):
        assert a['type'] == b['type'] and a['name'] == b['name'] and a['hypers'] == b['hypers']
        assert list(a['params']) == list(b['params'])
        for k in a['params']:
            assert a['params'][k].dtype == np.float32
            np.testing.assert_array_equal(a['params'][k], b['params'][k])
        for x, y in zip(a['comps'] + a['sinks'], b['comps'] + b['sinks']):
            same(x, y)
        if a['router'] is not None:
            same(a['router'], b['router'])
    ra, rb = serdes.encode_net(net), serdes.encode_net(net2)
    assert ra['type'] == rb['type'] and ra['hypers'] == rb['hypers']
    same(ra['root'], rb['root'])
    # parameter key order of the conv layer follows the reference (layer_types.py:174-179)
    keys = list(ra['root']['sinks'][0]['comps'][0]['params'])
    assert keys == ['w_horz_0', 'w_horz_1', 'w_horz_2', 'w_vert_0', 'w_vert_1', 'b_0', 'b_1', 'b_2']


def test_net_desc_schema_and_rendering():
    """nested dict {type, stats_tr, stats_ts, root{name, stats_*, sinks}} and the log text (desc.py:24-79)"""
    net = tiny_net('ac', k_cpt=1e-9)
    st = desc.state_tensors(net)
    leaves = list(net.leaves)
    assert (net, 'acc') in st and (net, 'moc') in st and (leaves[0], 'p_cor_by_cls') in st
    fake = {k: (0.5 if k[1] not in ('p_cor_by_cls', 'p_inc_by_cls') else [0.1] * 10) for k in st}
    d = {'type': 'ActorNet', 'stats_tr': {k: v for (t, k), v in fake.items() if t is net},
         'stats_ts': {k: v for (t, k), v in fake.items() if t is net},
         'root': desc.layer_desc(net.root, fake, fake)}
    assert d['root']['name'] == 'ToPyramid' and d['root']['sinks'][0]['name'] == 'ReConvMax'
    assert d['root']['sinks'][0]['sinks'][0]['name'] == 'LogReg'          # sinks[0] = classifier leaf
    text = desc.render_net_desc(d, 'nets/x/0000.npy — Epoch 1')
    assert text.startswith('┌') and 'Training Set:' in text and '[ActorNet] (acc=0.5; moc=0.5)' in text
    assert '↳ LogReg (c_err=0.5; p_cor=0.5; p_inc=0.5; p_tr=0.5)' in text


def test_dataset_schema_and_augmentation():
    ds = Dataset(archive=synthetic_archive(64, 32, (32, 32, 3), 10, seed=0), seed=1)
    assert ds.x0_shape == (32, 32, 3) and ds.y_shape == (10,)
    x, y = ds.augmented_training_batch(16)
    assert x.dtype == np.float64 and x.shape == (16, 32, 32, 3) and y.shape == (16, 10)
    sizes = [len(a) for a, _ in ds.training_set(24)]
    assert sizes == [24, 24, 16]                                        # ragged tail (data.py:42-47)


def test_cabi_exports_every_declared_symbol():
    protos = _cabi.parse_header()
    assert len(protos) >= 25 and 'mpnn_stencil_gemm' in protos and 'mpnn_route_fwd' in protos
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    assert _cabi.lib().mpnn_version() >= 100 and _cabi.lib().mpnn_has_umma() == 1


def test_no_cpu_fallback():
    """the product path refuses to run without a CUDA device"""
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    net = tiny_net('sr')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net.train.run({net.x0: np.zeros((2, 16, 16, 3), np.float32), net.y: np.eye(10, dtype=np.float32)[:2]})


@pytest.mark.parametrize('kind,prec', [('cnv', 'fp32'), ('cnv', 'bf16'), ('cnvpyr', 'bf16'), ('cnv', 'bf16x3'),
                                       ('cnv', 'bf16x6')])
def test_standalone_conv_chains_plan(kind, prec):
    """standalone Conv (layer_types.py:55-74) as tree nodes: Conv-BatchNorm-Rect on the image or on a pyramid
    scale picked by Select, classifier flattening the tensor; the plan uses the one-scale stage kernels"""
    from lib.engine import Engine
    net = tiny_net(kind, x0_shape=(16, 16, 1) if kind == 'cnvpyr' else (16, 16, 3))
    eng = Engine(net, precision=prec, impl=0 if prec == 'fp32' else 1, dry_run=True)
    plan = eng._plan(12, True, True)
    kinds = [getattr(op, 'kind', '') for op in plan.fwd_ops + plan.bwd_ops]
    n_fwd, n_wg, n_dg = {'bf16x3': (2, 6, 1), 'bf16x6': (4, 12, 2)}.get(prec, (2, 2, 1))
    assert kinds.count('conv_fwd') == n_fwd and kinds.count('conv_wgrad') == n_wg
    assert kinds.count('conv_dgrad') == n_dg               # none towards the input image
    assert eng.n_theta >= _n_params(net)
    rec = serdes.encode_net(net)
    assert [c['type'] for c in rec['root']['comps']][-3:] == ['Conv', 'BatchNorm', 'Rect'] or kind == 'cnvpyr'


@pytest.mark.parametrize('kind,prec,impl', [('sr', 'fp32', 0), ('ac', 'bf16', 1), ('cr', 'bf16', 1), ('actree', 'fp32', 0)])
def test_planner_dry_run(kind, prec, impl):
    from lib.engine import Engine
    net = tiny_net(kind, **({} if kind == 'sr' else dict(k_cpt=2e-9)))
    eng = Engine(net, precision=prec, impl=impl, dry_run=True)
    plan = eng._plan(12, True, True)
    assert 1 <= len(plan.pack_ops) <= 2 and plan.fwd_ops and plan.bwd_ops and len(plan.opt_ops) == 1
    assert eng.n_theta >= _n_params(net)
    kinds = [getattr(op, 'kind', '') for op in plan.fwd_ops + plan.bwd_ops]
    assert kinds.count('conv_fwd') == 6 * (2 if kind == 'actree' else 1) - (3 if kind == 'actree' else 0)
    assert kinds.count('conv_wgrad') == kinds.count('conv_fwd')
    with pytest.raises(RuntimeError, match='CUDA-only'):
        eng._run(plan.fwd_ops)
    with pytest.raises(NotImplementedError):
        from lib.layer_types import Chain, Rect
        from lib.net_types import SRNet
        Engine(SRNet(x0_shape=(16, 16, 3), y_shape=(10,), root=Chain(comps=[Rect()])), dry_run=True)


def test_checkpoint_format_roundtrip(tmp_path):
    """lib/checkpoint.py: the 'net' entry is exactly the write_net payload; step / RNG / momentum travel with it"""
    from lib import checkpoint, serdes
    from util import tiny_net
    net = tiny_net('ac', seed=3, k_cpt=4e-9)
    rng = np.random.default_rng(5)
    rng.random(7)
    path = str(tmp_path / 'ck.npy')
    checkpoint.save_checkpoint(path, net, step=1234, rng=rng)
    expect_next = rng.random(3)
    rec = checkpoint.net_record(path)
    ref = serdes.encode_net(net)
    assert rec['type'] == ref['type'] and rec['hypers'] == ref['hypers']
    flat = lambda r: [r['params'][k] for k in sorted(r['params'])] + [x for s in r.get('sinks', []) for x in flat(s)] \
        + [x for c in r.get('comps', []) for x in flat(c)] + (flat(r['router']) if r.get('router') else [])
    for a, b in zip(flat(rec['root']), flat(ref['root'])):
        np.testing.assert_array_equal(a, b)
    rng2 = np.random.default_rng(0)
    net2, step = checkpoint.load_checkpoint(path, rng=rng2)
    assert step == 1234 and type(net2).__name__ == 'ActorNet'
    np.testing.assert_array_equal(rng2.random(3), expect_next)        # sampling resumes where it stopped
    serdes.write_net(str(tmp_path / 'plain.npy'), net)
    assert checkpoint.net_record(str(tmp_path / 'plain.npy'))['type'] == 'ActorNet'


def test_net_file_cli_describe_and_roundtrip(tmp_path):
    """`python -m lib.checkpoint info|roundtrip`: read_net / write_net consistency on a plain file and a checkpoint"""
    from lib import checkpoint, serdes
    from util import tiny_net
    net = tiny_net('crtree', k_cpt=2e-9)
    plain, ck = str(tmp_path / 'n.npy'), str(tmp_path / 'c.npy')
    serdes.write_net(plain, net)
    checkpoint.save_checkpoint(ck, net, step=9)
    assert checkpoint.roundtrip(plain) == checkpoint.roundtrip(ck) > 20
    text = checkpoint.describe(ck)
    assert 'CriticNet' in text and 'step 9' in text and 'w_horz_0' in text and 'parameters' in text


def test_experiment_parallel_sharding(tmp_path):
    """one worker per GPU, each with its round-robin share of the experiment's nets and its own device"""
    import sys
    from lib.parallel import experiment_shards, run_experiment_parallel
    assert experiment_shards(range(8), [0, 1, 2]) == {0: [0, 3, 6], 1: [1, 4, 7], 2: [2, 5]}
    assert experiment_shards([4], [0, 1]) == {0: [4]}
    code = ("import os, sys; open(os.path.join(%r, os.environ['CUDA_VISIBLE_DEVICES']), 'w')"
            ".write(' '.join(sys.argv[1:]))" % str(tmp_path))
    rc = run_experiment_parallel(['-c', code, 'cifar10-ac', '--n-iter', '3'], [0, 1, 2], [5, 7], python=sys.executable)
    assert rc == 0
    assert (tmp_path / '5').read_text() == 'cifar10-ac --n-iter 3 --nets 0 2'
    assert (tmp_path / '7').read_text() == 'cifar10-ac --n-iter 3 --nets 1'
    assert run_experiment_parallel(['-c', 'raise SystemExit(3)'], [0, 1], [0, 1], python=sys.executable) == 3


def test_every_pdl_launched_kernel_waits_for_its_predecessor():
    """A kernel launched with the programmatic-stream-serialization attribute may start while the previous
    kernel of its stream is still running; it is only correct if it executes griddepcontrol.wait (SASS
    ACQBULK) before touching global memory.  Every kernel handed to mpnn_launch_pdl must contain it."""
    import glob
    import re
    import shutil
    import subprocess
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'multipath-nn_b200')
    so = os.path.join(root, 'lib', 'libmpnn_sm100.so')
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(so) or not os.path.exists(cuobjdump):
        pytest.skip('library or cuobjdump not available')
    launched = set()
    for f in glob.glob(os.path.join(root, 'csrc', '*.cu')):
        src = open(f).read()
        launched |= set(re.findall(r'mpnn_launch_pdl\(\s*(\w+)', src))
        if 'mpnn_launch_pdl(kern' in src:                       # function-pointer launch of the conv kernel
            launched.add('stencil_gemm_umma_kernel')
    launched.discard('kern')
    assert {'bn_relu_pool_fwd_kernel', 'bn_bwd_reduce_kernel', 'bn_relu_pool_bwd_kernel',
            'stencil_gemm_umma_kernel'} <= launched
    sass = subprocess.run([cuobjdump, '-sass', so], capture_output=True, text=True).stdout
    body, seen = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            seen = m.group(1)
            body[seen] = []
        elif seen is not None:
            body[seen].append(line)
    for kernel in launched:
        variants = [fn for fn in body if kernel in fn]
        assert variants, kernel
        for fn in variants:
            assert any('ACQBULK' in l for l in body[fn]), '%s is launched with PDL but never waits' % fn


def test_split_precision_passes_cover_exactly_the_products_kept():
    """The launch lists of the split-precision modes (lib/engine.py::_SPLIT_PASSES), restated in NumPy: bf16x3 keeps
    a_h*w_h + a_l*w_h + a_h*w_l of a two-way bf16 split (relative error ~2^-16), bf16x6 the six products of relative
    size >= 2^-16 of a three-way split (what is dropped is <= 2^-23) -- so a K = 1152 dot product (3x3 taps, 128
    channels) of fp32 operands comes out at fp32-accumulation accuracy in bf16x6 and at ~1e-5 in bf16x3."""
    from lib.engine import _SPLIT_PASSES

    def bf16(a):                       # round-to-nearest-even to 8 significant bits, as __float2bfloat16_rn
        u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return u.astype(np.uint32).view(np.float32)

    def parts(x, n):                   # mpnn_split_planes / mpnn_split_planes3, pack modes 0 / 4 / 8
        out, r = [], np.asarray(x, np.float32)
        for _ in range(n):
            h = bf16(r)
            out.append(h)
            r = (r - h).astype(np.float32)
        return out
    rng = np.random.default_rng(5)
    a = rng.standard_normal((64, 1152)).astype(np.float32)
    w = (rng.standard_normal((1152, 32)) / np.sqrt(1152)).astype(np.float32)
    exact = a.astype(np.float64) @ w.astype(np.float64)
    mode_part = {0: 0, 4: 1, 8: 2}
    errs = {}
    for n, passes in _SPLIT_PASSES.items():
        ap, wp = parts(a, n), parts(w, n)
        assert np.abs(sum(p.astype(np.float64) for p in ap) - a).max() <= 2.0 ** (-8 * n) * np.abs(a).max()
        acc, kept = np.zeros_like(exact), set()
        for na0, na1, wmodes in passes:
            a_idx = list(range(na0)) + list(range(na1))          # A0 = the first na0 parts, A1 = the first na1
            assert len(a_idx) == len(wmodes) == 3
            for i, m in zip(a_idx, wmodes):
                kept.add((i, mode_part[m]))
                acc += ap[i].astype(np.float64) @ wp[mode_part[m]].astype(np.float64)
        assert kept == {(i, j) for i in range(n) for j in range(n) if i + j < n}      # the pairs the weight gradient uses too
        errs[n] = float(np.linalg.norm(acc - exact) / np.linalg.norm(exact))
    assert 1e-7 < errs[2] < 3e-5 and errs[3] < 3e-7, errs


def test_small_batch_plan_choices():
    """What the planner picks at the reference's batch and above it (dry run; arch_and_hypers.py:19-35):
    the coarsest scale of every stage takes the one-launch BatchNorm kernels while its tensor has at most 8192 pixels
    (B <= 512 at 4x4) and the two-pass pair beyond; the router tails wait for the head GEMMs that feed a router only."""
    import arch_and_hypers as ah
    from lib import layer_types
    from lib.engine import Engine
    layer_types.seed(0)
    net = ah.ac_chain(k_cpt=4e-9)((32, 32, 3), (10,))
    eng = Engine(net, precision='bf16', impl=1, dry_run=True)
    for B, small in ((128, True), (512, True), (1024, False)):
        plan = eng._plan(B, True, True)
        ops = plan.fwd_ops + plan.bwd_ops
        one = [op for op in ops if getattr(op, 'kind', '') == 'bn_bwd' and '1-launch' in getattr(op, 'desc', '')]
        red_h4 = [op for op in ops if getattr(op, 'kind', '') == 'bn_bwd_reduce' and getattr(op, 'desc', '').startswith('H4 ')]
        assert (len(one) == 8 and not red_h4) if small else (not one and red_h4), (B, len(one), len(red_h4))
        assert all(op.desc.startswith('H4 ') for op in one)
    plan = eng._plan(128, True, True)
    tails = [op for op in plan.fwd_ops if getattr(op, 'deps', None) and all(getattr(d, 'kind', '') == 'fc_fwd' for d in op.deps)
             and len(op.deps) >= 2 and getattr(op, 'lane', 0) == 0]
    assert len(tails) == 1 and len(tails[0].deps) == 7 and all(d.feeds_router for d in tails[0].deps)     # 8 stages, 7 routers
    heads = [op for op in plan.fwd_ops if getattr(op, 'kind', '') == 'fc_fwd']
    assert len(heads) == 8 and sum(1 for h in heads if not h.feeds_router) == 1
