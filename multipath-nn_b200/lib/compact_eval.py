"""Compacted 'ev'-mode evaluation: the second hot loop of the reference
(/root/reference/scripts/lib/desc.py:10-22 called from scripts/train-nets:144-155 -- full passes over the
training and test sets every t_log steps).

The reference evaluates every node of the tree on the whole batch and multiplies the statistics with the
one-hot p_ev (scripts/lib/net_types.py:127-131, scripts/train-nets:117-130).  In 'ev' mode BatchNorm uses
its running moments, so examples are independent and an example only needs the nodes ON ITS OWN PATH:
after every switch the batch is compacted (mpnn_route_compact: first-max argmax, ballot + prefix scan into
per-sink index lists), the activation pyramids of the examples that continue are gathered into the child's
input buffers (mpnn_gather_images) and the child runs on the smaller batch; classifiers are scored only on
the examples that exit at them (mpnn_leaf_stats).  All sums stay on the device and are read back once per
data set.

Exact for every p_ev-weighted statistic -- acc, moc, and per leaf p_cor, p_inc, p_cor_by_cls, p_inc_by_cls
(what make-acc-eff-plots / make-nlds / make-routing-hists read).  The three entries of `state_tensors` that
are NOT p_ev-weighted (per-leaf c_err and p_tr, per-switch x_rte; train-nets:125-128) are defined on examples
a compacted pass never evaluates: they are absent from its result (use the dense `Engine.eval_stats` for them).
"""
import ctypes

import numpy as np
import torch

from lib.engine import BF16, _RT_FWD, _ru, _vp

__all__ = ['CompactEvaluator']


class CompactEvaluator:
    def __init__(self, eng, batch=4096):
        if not eng.dynamic:
            raise ValueError('compacted evaluation is for dynamically-routed nets (an SRNet has one path)')
        if eng.split:
            raise NotImplementedError('compacted evaluation runs in fp32 or bf16 precision (not the split modes bf16x3 / bf16x6)')
        if any(getattr(nd, 'maxpool', False) or getattr(nd, 'gmp', False) for nd in eng.nodes):
            raise NotImplementedError('compacted evaluation does not cover MaxPool / GlobalMaxPool blocks')
        if any(nd.loss != 'ce' for nd in eng.regs):
            raise NotImplementedError('compacted evaluation scores Softmax + CrossEntropyError classifiers only')
        self.eng, self.L, self.B = eng, eng.L, int(batch)
        self.plan = plan = eng._plan(self.B, False, False)      # buffers, packed operands, BN constants
        self.n_cls = eng.net.hypers.y_shape[0]
        B = self.B
        dev = eng.dev
        zi = lambda *shape: torch.zeros(shape, dtype=torch.int32, device=dev)
        self.sw = {}
        for nd in eng.nodes:
            if len(nd.kids) > 1:
                ns = len(nd.kids)
                rt = plan.rtr[nd.idx]
                P = lambda lay, k: eng.tptr(getattr(lay.params, k))
                tab = plan._desc_table(_RT_FWD, [dict(
                    Z1=rt.Z1, g1=P(rt.bn1, 'γ'), b1=P(rt.bn1, 'β'), m1=P(rt.bn1, 'm_avg'), v1=P(rt.bn1, 'v_avg'),
                    W2=P(rt.fc2, 'w'), bias2=P(rt.fc2, 'b'), g2=P(rt.bn2, 'γ'), b2=P(rt.bn2, 'β'),
                    m2=P(rt.bn2, 'm_avg'), v2=P(rt.bn2, 'v_avg'), W3=P(rt.fc3, 'w'), bias3=P(rt.fc3, 'b'),
                    Z2=rt.Z2, R=rt.R, save=rt.save, ns=ns)])
                self.sw[nd.idx] = dict(tab=tab, pos=zi(ns, B), orig=zi(ns, B), count=zi(ns),
                                       host=torch.zeros(ns, dtype=torch.int32).pin_memory())
        # compacted inputs of every conv stage below a switch: one planes tensor per scale it consumes
        self.cin = {}
        for nd in eng.nodes:
            if nd.kind == 'rcm' and len(eng.nodes[nd.parent].kids) > 1:
                st = plan.node[nd.idx]
                self.cin[nd.idx] = [plan.planes(s.C, s.geo) for s in st.pin]
        n_leaf = len(eng.regs)
        self.acc = torch.zeros((n_leaf, 2 + 2 * self.n_cls), dtype=torch.float64, device=dev)
        self.visits = np.zeros(len(eng.nodes), np.int64)
        self.n_seen = 0
        self._prepared = False

    # ------------------------------------------------------------------ #
    def reset(self):
        self.acc.zero_()
        self.visits[:] = 0
        self.n_seen = 0
        self._prepared = False            # parameters may have changed since the last data set

    def _S(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.eng.dev).cuda_stream)

    def _prepare(self):
        """once per data set: pack the conv / head operands and turn the running BN moments into scale / shift"""
        eng, plan = self.eng, self.plan
        eng.stream = self._S()
        for op in plan.pack_ops:
            op()
        for op in plan.fwd_ops:
            if getattr(op, 'kind', '') == 'bn_finalize':
                op()
        self._prepared = True

    # ------------------------------------------------------------------ #
    def run_batch(self, x0, y, k_cpt=None):
        """accumulate the statistics of one batch (n <= batch examples; host or device arrays)"""
        eng, plan, L = self.eng, self.plan, self.L
        with torch.cuda.device(eng.dev):
            if not self._prepared:
                self._prepare()
            n = int(x0.shape[0])
            if n > self.B:
                raise ValueError('batch %d > evaluator capacity %d' % (n, self.B))
            hy = eng.net.hypers
            x0 = eng._to_dev(x0, (n,) + tuple(hy.x0_shape), 'x0')
            y = eng._to_dev(y, (n,) + tuple(hy.y_shape), 'y')
            plan.x0[:n].copy_(x0, non_blocking=True)
            plan.y[:n].copy_(y, non_blocking=True)
            if hy.dyn_k_cpt:
                kc = np.unique(np.asarray(k_cpt, dtype=np.float32).reshape(-1))
                if kc.size != 1:
                    raise NotImplementedError('compacted evaluation needs one k_cpt for the whole batch '
                                              '(the length-1 feed of train-adaptive-nets:102-105)')
                plan.kextra.fill_(float(kc[0] * np.float32(hy.α_cpt)))
                for col in plan.kplanes:
                    col.fill_(float(kc[0] * np.float32(hy.α_cpt)))
            self.n_seen += n
            root = eng.nodes[0]
            st = plan.node[0]
            H0, W0, C0 = hy.x0_shape
            for i, slot in enumerate(st.out):
                L.pack_input(_vp(plan.x0), n, H0, W0, C0, 2 ** i, _vp(slot.t), min(slot.C, (C0 + 7) // 8 * 8), slot.geo.G, slot.geo.P,
                             eng.dtype, self._S())
            self.visits[0] += n
            for k in root.kids:
                self._node(eng.nodes[k], n, None, None, None)

    # ------------------------------------------------------------------ #
    def _node(self, nd, n, orig, pos, count_dev):
        """evaluate node `nd` on the n examples routed to it.  orig: device list of their original ids
        (None = identity); pos / count_dev: their rows in the parent's compact batch (None = all of them)."""
        eng, plan, L = self.eng, self.plan, self.L
        if n == 0:
            return
        self.visits[nd.idx] += n
        if nd.kind == 'reg':
            par = eng.nodes[nd.parent]
            r = plan.reg[nd.idx]
            if not plan.umma_heads:
                pst = plan.node[par.idx]
                # logits of the parent's compact batch (rows of the examples that exit here are read through pos)
                L.fc_fwd(_vp(pst.feat), pst.F, plan.Balloc, self._n_parent, eng.tptr(r.fc.params.w),
                         eng.tptr(r.fc.params.b), None, self.n_cls, _vp(r.Zbuf), eng.dtype, self._S())
            L.leaf_stats(_vp(r.Zbuf), r.ldz, self.n_cls, _vp(plan.y), _vp(pos), _vp(orig), _vp(count_dev), n,
                         _vp(self.acc[nd.err]), self._S())
            return
        st = plan.node[nd.idx]
        dt, impl = eng.dtype, eng.impl
        ins = [s.t for s in st.pin]
        if pos is not None:                         # below a switch: gather the inputs of the examples that continue
            for k, (slot, dst) in enumerate(zip(st.pin, self.cin[nd.idx])):
                g = slot.geo
                L.gather_images(_vp(slot.t), self._n_parent, g.P, _vp(pos), _vp(count_dev), _vp(dst), n, g.P,
                                slot.C, g.H, g.W, g.G, dt, self._S())
            ins = self.cin[nd.idx]
        for k, sc in enumerate(st.sc):
            g = sc.geo
            prev = st.sc[k - 1] if k > 0 else None
            L.stencil_gemm(_vp(ins[k]), sc.K0, _vp(prev.pooled) if prev is not None else None, sc.K1, _vp(sc.Wf), 9,
                           eng.tptr(sc.bk), _vp(sc.lin), sc.N, 0, None, 0, 0, n, g.H, g.W, g.G, g.P,
                           None, 0, None, dt, dt, impl, self._S())
            if sc.live or sc.pooled is not None:
                L.bn_relu_pool_fwd(_vp(sc.lin), sc.N, n, g.H, g.W, g.G, g.P, _vp(sc.ss) if sc.live else None,
                                   _vp(sc.act), _vp(sc.pooled), sc.geo_p.P if sc.pooled is not None else 0,
                                   _vp(sc.feat), plan.Balloc, dt, self._S())
        if not nd.kids:
            return
        dyn_k = bool(eng.net.hypers.dyn_k_cpt)
        rt = plan.rtr.get(nd.idx)
        if plan.umma_heads and nd.idx in plan.heads:
            hd = plan.heads[nd.idx]
            outs = ([(hd.Z16, 16)] if hd.leaf_off is not None else []) + ([(rt.Z1, 16)] if rt is not None else [])
            outs.append((None, 0))
            (o0, n0), (o1, n1) = outs[0], outs[1]
            L.stencil_gemm(_vp(st.feat), st.Fext, None, 0, _vp(hd.Wfc), 1, _vp(hd.bias), _vp(o0), n0, 0, _vp(o1), n1, 0,
                           n, 0, 0, 0, plan.Balloc, None, 0, None, BF16, 2, 1, self._S())
        elif rt is not None:
            L.fc_fwd(_vp(st.feat), st.F, plan.Balloc, n, eng.tptr(rt.fc1.params.w), eng.tptr(rt.fc1.params.b),
                     _vp(plan.kextra) if dyn_k else None, 16, _vp(rt.Z1), dt, self._S())
        if len(nd.kids) == 1:                       # no switch: the only sink sees every example of this node
            self._n_parent = n
            self._node(eng.nodes[nd.kids[0]], n, orig, None, None)
            return
        sw = self.sw[nd.idx]
        ns = len(nd.kids)
        L.router_tail_fwd_batched(_vp(sw['tab']), 1, n, 16, float(rt.bn1.hypers.d), float(rt.bn1.hypers.ε), 0, self._S())
        L.route_compact(_vp(rt.R), ns, ns, n, _vp(orig), self.B, None, _vp(sw['pos']), _vp(sw['orig']),
                        _vp(sw['count']), self._S())
        sw['host'].copy_(sw['count'], non_blocking=True)
        torch.cuda.current_stream(eng.dev).synchronize()            # the child batches are launch parameters
        counts = [int(c) for c in sw['host']]
        for s, kid in enumerate(nd.kids):
            self._n_parent = n
            self._node(eng.nodes[kid], counts[s], sw['orig'][s], sw['pos'][s], sw['count'][s:s + 1])

    # ------------------------------------------------------------------ #
    def result(self):
        """{(object, name): mean over the data set} in `mean_net_state`'s form (desc.py:10-22) for the exact keys"""
        eng = self.eng
        n = max(self.n_seen, 1)
        a = self.acc.cpu().numpy() / n                       # the one device-to-host read of the data set
        nc = self.n_cls
        out = {}
        acc = 0.0
        for nd in eng.regs:
            row = a[nd.err]
            l = nd.layer
            out[(l, 'p_cor')] = float(row[0])
            out[(l, 'p_inc')] = float(row[1])
            out[(l, 'p_cor_by_cls')] = row[2:2 + nc].tolist()
            out[(l, 'p_inc_by_cls')] = row[2 + nc:2 + 2 * nc].tolist()
            acc += float(row[0])
        moc = 0.0
        for nd in eng.nodes:
            ops = nd.layer.n_ops + (nd.router.n_ops if nd.router is not None else 0)
            moc += self.visits[nd.idx] * float(ops)
        out[(eng.net, 'acc')] = acc
        out[(eng.net, 'moc')] = moc / n
        return out
