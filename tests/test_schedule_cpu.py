"""Static race check of the step scheduler (CPU, no GPU): every op of a plan is executed against a
RECORDING stand-in for the C ABI, which yields the device buffers it reads and writes (const / non-const
pointer parameters of include/mpnn.h, plus the pointers inside descriptor tables and fused-BN structs).
Two ops of one list that touch the same buffer, at least one writing, must be ordered by the scheduler's
happens-before relation: same lane (stream order) or a chain of explicit `deps` (events).  This is the
property `Engine._run` relies on; a forgotten dependency shows up here, not as a flaky GPU result."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

from lib import _cabi, engine as E  # noqa: E402
from util import tiny_net  # noqa: E402


def _roles():
    """{function: [(argname, 'r' | 'w' | None)]} from the header: const pointers are read, others written"""
    src = open(_cabi.HEADER).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\bint\s+(mpnn_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        roles = []
        for a in ' '.join(m.group(2).split()).split(','):
            a = a.strip()
            name = re.findall(r'\w+', a)[-1] if a and a != 'void' else ''
            roles.append((name, None if '*' not in a or name in ('stream', 'n_parts') else ('r' if 'const' in a else 'w')))
        out[m.group(1)[5:]] = roles
    return out


# pointers hidden inside structs: field -> role ('r', 'w', 'rw')
_STRUCTS = {
    'router_tail_fwd_batched': (E._RT_FWD, dict(Z1='r', g1='r', b1='r', m1='rw', v1='rw', W2='r', bias2='r', g2='r', b2='r',
                                               m2='rw', v2='rw', W3='r', bias3='r', Z2='w', R='w', save='w')),
    'router_tail_bwd_batched': (E._RT_BWD, dict(Z1='r', Z2='r', dR='r', g1='r', b1='r', W2='r', g2='r', b2='r', W3='r',
                                               save='r', dg1='w', dbt1='w', dW2='w', dbias2='w', dg2='w', dbt2='w',
                                               dW3='w', dbias3='w', dZ1='w', scratch='w', dZ1p='w', dbias1='w')),
    'pack_weights_batched': (E._PACK, dict(w='r', packed='w')),
}
_HOST_STRUCTS = {
    'conv_bn_stats': ('bn', E._BN_FUSE, dict(acc='rw', gamma='r', beta='r', m_avg='rw', v_avg='rw', ss='w', mr='w')),
    'conv_acc_bn_stats': ('bn', E._BN_FUSE, dict(acc='rw', gamma='r', beta='r', m_avg='rw', v_avg='rw', ss='w', mr='w')),
    'bn_relu_pool_fwd_acc': ('bn', E._BN_FUSE, dict(acc='r', gamma='r', beta='r', m_avg='rw', v_avg='rw', ss='w', mr='w')),
    'bn_bwd_reduce_fused': ('f', E._BN_BWD_FUSE, dict(acc='rw', sums='w', dgamma='w', dbeta='w')),
    'conv_dgrad_bn_reduce': ('epi', E._BN_BWD_EPI, dict(lin='r', ss='r', mr='r', acc='rw', sums='w', dgamma='w',
                                                        dbeta='w')),
}


class Recorder:
    def __init__(self, real, tensors):
        self.protos, self.launches, self.calls, self.roles, self.tensors = real.protos, 0, [], _roles(), tensors

    def __getattr__(self, name):
        if name.startswith('mpnn_'):
            name = name[5:]

        def call(*args):
            reads, writes = set(), set()

            def add(ptr, role):
                if ptr:
                    if 'r' in role:
                        reads.add(ptr)
                    if 'w' in role:
                        writes.add(ptr)
            for (argname, role), v in zip(self.roles[name], args):
                if role is None or not isinstance(v, ctypes.c_void_p) or not v.value:
                    continue
                if name in _STRUCTS and argname == 'descs':
                    dt, fields = _STRUCTS[name]
                    tab = self.tensors(v.value)
                    for row in tab.numpy().view(dt):
                        for f, r in fields.items():
                            add(int(row[f]), r)
                elif name in _HOST_STRUCTS and argname == _HOST_STRUCTS[name][0]:
                    _, dt, fields = _HOST_STRUCTS[name]
                    row = np.ctypeslib.as_array(ctypes.cast(v, ctypes.POINTER(ctypes.c_uint8)), (dt.itemsize,)).view(dt)[0]
                    for f, r in fields.items():
                        add(int(row[f]), r)
                else:
                    add(v.value, role)
            self.calls.append((name, reads, writes))
            self.launches += 1
            return 0
        return call


def _tensor_index(*roots):
    """all torch tensors reachable from the given objects -> lookup(pointer) -> containing tensor"""
    seen, found = set(), {}
    for r in roots:                              # the zero-filled arenas a plan's buffers are carved from are not buffers
        for a in list(getattr(r, 'arenas', [])) + [getattr(r, '_accpool', None)]:   # (nor is the pool the deferred
            seen.add(id(a))                                                          #  BN accumulators are slices of)

    def walk(o, depth=0):
        if id(o) in seen or depth > 6:
            return
        seen.add(id(o))
        if isinstance(o, torch.Tensor):
            if o.numel():                     # (an arena and its first view share a pointer: the view is the buffer)
                old = found.get(o.data_ptr())
                if old is None or o.numel() * o.element_size() < old.numel() * old.element_size():
                    found[o.data_ptr()] = o
        elif isinstance(o, dict):
            for v in o.values():
                walk(v, depth + 1)
        elif isinstance(o, (list, tuple)):
            for v in o:
                walk(v, depth + 1)
        elif hasattr(o, '__dict__') and type(o).__module__ in ('types', 'lib.engine'):
            for v in vars(o).values():
                walk(v, depth + 1)
    for r in roots:
        walk(r)
    spans = sorted((p, p + t.numel() * t.element_size(), t) for p, t in found.items())

    def lookup(ptr):
        best = None
        for lo, hi, t in spans:
            if lo <= ptr < hi and (best is None or hi - lo < best[1] - best[0]):
                best = (lo, hi, t)            # innermost view wins (tables live in their own tensors)
        return best
    return lookup


def _check(eng, plan):
    flat = {}
    for t in (eng.theta, eng.grad, eng.state, eng.accum):
        flat[t.data_ptr()] = t.data_ptr() + t.numel() * 4
    lookup = _tensor_index(eng, plan)

    def ident(ptr):
        for lo, hi in flat.items():
            if lo <= ptr < hi:
                return ptr                 # parameter / gradient segments: exact pointer
        hit = lookup(ptr)
        return hit[0] if hit else ptr      # everything else: the containing allocation
    problems = []
    for lname in ('pack_ops', 'fwd_ops', 'bwd_ops', 'opt_ops'):
        ops = getattr(plan, lname)
        acc = []
        for op in ops:
            eng.L.calls = []
            op()
            r = {ident(p) for _, rr, _ in eng.L.calls for p in rr}
            w = {ident(p) for _, _, ww in eng.L.calls for p in ww}
            acc.append((op, r, w, [c[0] for c in eng.L.calls]))
        n = len(ops)
        index = {id(op): i for i, op in enumerate(ops)}
        before = [set() for _ in range(n)]          # happens-before closure: before[j] = ops ordered before j
        last_in_lane = {}
        for j, op in enumerate(ops):
            preds = [index[id(d)] for d in getattr(op, 'deps', ()) if id(d) in index]
            lane = getattr(op, 'lane', 0)
            if lane in last_in_lane:
                preds.append(last_in_lane[lane])
            if getattr(op, 'wait_all', False):         # Engine._run orders it after everything issued so far
                preds.extend(range(j))
            for i in preds:
                assert i < j, 'dependency on a later op'
                before[j] |= before[i] | {i}
            last_in_lane[lane] = j
        for j in range(n):
            _, rj, wj, nj = acc[j]
            for i in range(j):
                _, ri, wi, ni = acc[i]
                clash = (wi & (rj | wj)) | (ri & wj)
                if clash and i not in before[j]:
                    problems.append((lname, i, ni, getattr(ops[i], 'lane', 0), j, nj, getattr(ops[j], 'lane', 0)))
    return problems


@pytest.mark.parametrize('prec,impl', [('bf16', 1), ('fp32', 0), ('bf16x3', 1), ('bf16x6', 1)])
@pytest.mark.parametrize('kind,hy', [('sr', {}), ('ac', dict(k_cpt=4e-9)), ('cr', dict(k_cpt=4e-9)),
                                     ('actree', dict(k_cpt=2e-9)), ('ac', dict(dyn_k_cpt=True)), ('cnv', {}),
                                     ('cnvmp', {}), ('cnvgmp', {}), ('cnvact', {}), ('cnvdrop', {}), ('aclln', dict(k_cpt=4e-9)), ('acsce', dict(k_cpt=4e-9)), ('srsq', {})])
def test_every_conflicting_pair_of_launches_is_ordered(kind, hy, prec, impl):
    if prec in ('bf16x3', 'bf16x6') and kind in ('cnvmp', 'cnvgmp', 'cnvact', 'cnvdrop'):
        pytest.skip('MaxPool / GlobalMaxPool / ActivityError blocks are not served in the split-precision modes')
    net = tiny_net(kind, **hy)
    eng = E.Engine(net, precision=prec, impl=impl, dry_run=True)
    holder = {}
    eng.L = Recorder(eng.L, lambda ptr: holder['lookup'](ptr)[2])
    plan = eng._plan(12, True, True)
    holder['lookup'] = _tensor_index(eng, plan)
    problems = _check(eng, plan)
    assert not problems, problems[:8]


def test_the_check_sees_a_missing_dependency():
    """fault injection: without its event edge a weight-gradient launch races the BN backward that feeds it"""
    net = tiny_net('ac', k_cpt=4e-9)
    eng = E.Engine(net, precision='bf16', impl=1, dry_run=True)
    holder = {}
    eng.L = Recorder(eng.L, lambda ptr: holder['lookup'](ptr)[2])
    plan = eng._plan(12, True, True)
    holder['lookup'] = _tensor_index(eng, plan)
    assert not _check(eng, plan)
    victim = next(op for op in plan.bwd_ops if getattr(op, 'kind', '') == 'conv_wgrad')
    victim.deps = []
    problems = _check(eng, plan)
    assert problems and all('stencil_wgrad' in p[5] for p in problems), problems
    for op in plan.fwd_ops:
        if hasattr(op, 'deps'):
            op.deps = []
    assert any(p[0] == 'fwd_ops' for p in _check(eng, plan))


@pytest.mark.parametrize('which', ['sr_chain', 'ac_chain', 'cr_chain', 'ac_tree', 'cr_tree_dyn'])
def test_reference_architectures_are_race_free(which):
    """the same check on the nets of arch_and_hypers.py (8 stages, 4 pyramid scales; trees have 2-3 sinks)"""
    import arch_and_hypers as ah
    from lib import layer_types
    layer_types.seed(0)
    make = {'sr_chain': lambda: ah.sr_chain(8), 'ac_chain': lambda: ah.ac_chain(k_cpt=4e-9),
            'cr_chain': lambda: ah.cr_chain(k_cpt=4e-9, optimistic=True), 'ac_tree': lambda: ah.ac_tree(k_cpt=4e-9),
            'cr_tree_dyn': lambda: ah.cr_tree(dyn_k_cpt=True)}[which]
    net = make()((32, 32, 3), (10,))
    eng = E.Engine(net, precision='bf16', impl=1, dry_run=True)
    eng.fuse_bn_red_min_rows = 0          # exercise the data gradients that carry the BN-backward sums (large batches)
    holder = {}
    eng.L = Recorder(eng.L, lambda ptr: holder['lookup'](ptr)[2])
    plan = eng._plan(8, True, True)
    holder['lookup'] = _tensor_index(eng, plan)
    problems = _check(eng, plan)
    assert not problems, problems[:8]
    if which != 'ac_tree' and which != 'cr_tree_dyn':
        assert any('bnred' in getattr(op, 'desc', '') for op in plan.bwd_ops)
    ev = eng._plan(8, False, False)                  # inference plan (running BN moments, no backward)
    holder['lookup'] = _tensor_index(eng, ev)
    problems = _check(eng, ev)
    assert not problems, problems[:8]
