"""Independent NumPy fp64 forward restatement (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED (see oracle/torch_ref.py).  This twin exists so that the
PyTorch oracle is not checked only against itself: it re-derives the forward
value of every hot-path layer, the routing products and c_tot with plain
NumPy (shift-and-einsum convolution, reshape pooling, explicit softmax),
sharing no code with torch_ref.py.  tests/test_oracle.py compares the two and
uses this forward for finite-difference gradient checks.

Only the layer types the configured architectures use are covered
(/root/reference/scripts/arch_and_hypers.py:45-70): Chain, ToPyramid,
MultiscaleConvMax, MultiscaleBatchNorm, MultiscaleRect, Select, LinTrans,
BatchNorm, Rect, Softmax, CrossEntropyError.
"""
import numpy as np

__all__ = ['forward']


def _eps(hy, default=1e-6):
    """The reference spells the hyper `\u03f5` (U+03F5); Python NFKC-normalises
    identifiers, so the attribute / record key is `\u03b5` (U+03B5)."""
    for k in ('\u03b5', '\u03f5'):
        if k in hy:
            return hy[k]
    return default


def conv3_same(x, w):
    """NHWC x HWIO, stride 1, SAME, cross-correlation (layer_types.py:106-107)."""
    kh, kw = w.shape[:2]
    B, H, W, _ = x.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((B, H + kh - 1, W + kw - 1, x.shape[3]))
    xp[:, pt:pt + H, pl:pl + W] = x
    out = np.zeros((B, H, W, w.shape[3]))
    for i in range(kh):
        for j in range(kw):
            out += np.einsum('bhwc,cd->bhwd', xp[:, i:i + H, j:j + W], w[i, j])
    return out


def pool2(x):
    """2x2/2 max pool on even sizes (layer_types.py:109-110)."""
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max((2, 4))


def bn(p, hy, x, mode):
    """layer_types.py:219-239 (EMA side effect not modelled here)."""
    eps = _eps(hy)
    g, b = np.float64(p['γ']), np.float64(p['β'])
    if mode == 'tr':
        ax = tuple(range(x.ndim - 1))
        m, v = x.mean(ax), x.var(ax)
    else:
        m, v = np.float64(p['m_avg']), np.float64(p['v_avg'])
    return g * (x - m) / np.sqrt(v + eps) + b


def softmax(z):
    e = np.exp(z - z.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True)


def link(rec, x, y, mode):
    """-> dict(x, c_err, c_mod, n_ops, d_cor)."""
    t, hy, p = rec['type'], rec['hypers'], rec['params']
    r = dict(x=x, c_err=0.0, c_mod=0.0, n_ops=0, d_cor=None)
    if t == 'Chain':
        for c in rec['comps']:
            o = link(c, x, y, mode)
            x = o['x']
            r['c_err'] = r['c_err'] + o['c_err']
            r['c_mod'] = r['c_mod'] + o['c_mod']
            r['n_ops'] += o['n_ops']
            r['d_cor'] = o['d_cor']
        r['x'] = x
    elif t == 'ToPyramid':
        r['x'] = [x[:, ::2 ** i, ::2 ** i] for i in range(hy['n_scales'])]
    elif t == 'MultiscaleConvMax':
        n = len(hy['n_chan'])
        xin = x[-n:]
        outs = []
        for k in range(n):
            wh = np.float64(p['w_horz_%i' % k])
            o = np.float64(p['b_%i' % k]) + conv3_same(xin[k], wh)
            r['c_mod'] += hy['k_l2'] * (wh ** 2).sum()
            per_px = wh.size
            if k:
                wv = np.float64(p['w_vert_%i' % (k - 1)])
                o = o + conv3_same(pool2(outs[-1]), wv)
                r['c_mod'] += hy['k_l2'] * (wv ** 2).sum()
                per_px += wv.size
            r['n_ops'] += o.shape[1] * o.shape[2] * per_px
            outs.append(o)
        r['x'] = outs
    elif t == 'MultiscaleBatchNorm':
        r['x'] = [bn(c['params'], c['hypers'], v, mode) for c, v in zip(rec['comps'], x)]
    elif t == 'MultiscaleRect':
        r['x'] = [np.maximum(v, 0) for v in x]
    elif t == 'Rect':
        r['x'] = np.maximum(x, 0)
    elif t == 'Select':
        r['x'] = x[hy.get('i', 0)]
    elif t == 'BatchNorm':
        r['x'] = bn(p, hy, x, mode)
    elif t == 'LinTrans':
        w = np.float64(p['w'])
        r['x'] = x.reshape(len(x), -1) @ w + np.float64(p['b'])
        r['c_mod'] = hy.get('k_l2', 0) * (w ** 2).sum()
        r['n_ops'] = w.shape[0] * w.shape[1]
    elif t == 'Softmax':
        r['x'] = softmax(x)
    elif t == 'CrossEntropyError':
        e = _eps(hy)
        pc = e / y.shape[1] + (1 - e) * x
        r['c_err'] = -(y * np.log(pc)).sum(1)
        r['d_cor'] = (x.argmax(1) == y.argmax(1)).astype(np.float64)
    else:
        raise NotImplementedError(t)
    return r


def _leaves(rec):
    return 1 if not rec['sinks'] else sum(map(_leaves, rec['sinks']))


def forward(record, x0, y, mode='ev', tau=None, k_cpt=None, frozen=None):
    """-> dict(c_tot, nodes{path: dict(p_tr, p_ev, c_err, d_cor, r, n_ops)}).
    Follows net_types.py:85-97 (SR), :103-181 (actor), :187-284 (critic).

    `frozen`: the `nodes` dict of a previous forward; its p_tr / c_ev / c_opt are
    used wherever the reference applies tf.stop_gradient, so that finite
    differences of c_tot reproduce the gradient TF would compute."""
    hy = record['hypers']
    kind = record['type']
    x0 = np.float64(x0); y = np.float64(y)
    B = len(x0)
    dyn_k = kind != 'SRNet' and hy.get('dyn_k_cpt', False)
    if kind != 'SRNet':
        tau = hy['τ'] if tau is None else tau
        eps = _eps(hy)
        k_cpt = (np.float64(k_cpt).reshape(-1) if dyn_k
                 else (hy.get('k_cpt', 0.0) if k_cpt is None else k_cpt))
    nodes = {}
    order = []

    def walk(rec, x, path):
        o = link(rec, x, y, mode)
        o['rec'] = rec
        o['router'] = None
        if rec['router'] is not None:
            xr = o['x']
            if dyn_k:
                f = lambda v: np.concatenate(
                    [v.reshape(B, -1), hy.get('α_cpt', 1e7) * k_cpt[:, None] * np.ones((B, 1))], 1)
                xr = [f(v) for v in xr] if isinstance(xr, list) else f(xr)
            o['router'] = link(rec['router'], xr, y, mode)
        nodes[path] = o
        order.append(path)
        for i, s in enumerate(rec['sinks']):
            walk(s, o['x'], (path + '/' if path else '') + str(i))

    walk(record['root'], x0, '')
    one = np.ones(B)
    if kind == 'SRNet':
        tot = sum(np.mean(o['c_err'] * one) + o['c_mod'] for o in nodes.values())
        for o in nodes.values():
            o['p_ev'] = one; o['p_tr'] = None
        return dict(c_tot=tot, nodes=nodes, order=order)
    n_root = _leaves(record['root'])
    critic = kind == 'CriticNet'

    def route(path, p_tr, p_ev):
        o = nodes[path]
        rec = o['rec']
        o['p_tr'], o['p_ev'] = p_tr, p_ev
        kids = [(path + '/' if path else '') + str(i) for i in range(len(rec['sinks']))]
        ce = o['c_err']
        if critic and hy.get('use_cls_err', False):
            ce = 1 - (o['d_cor'] if o['d_cor'] is not None else 1)
        if len(kids) < 2:
            for k in kids:
                route(k, p_tr, p_ev)
            if critic:
                base = ce + k_cpt * o['n_ops']
                o['c_ev'] = base + sum(nodes[k]['c_ev'] for k in kids)
                o['c_opt'] = base + sum(nodes[k]['c_opt'] for k in kids)
                o['c_cre'] = 0.0
            return
        r = o['router']['x']
        e_here = eps * _leaves(rec) / n_root
        e_kids = np.array([eps * _leaves(s) / n_root for s in rec['sinks']])
        sm = softmax(r / tau)
        child_p = (p_tr[:, None] - e_here) * sm + e_kids       # = p_tr * pi_tr
        dec = r.argmax(1)                                      # first max
        o['dec'] = dec
        for i, k in enumerate(kids):
            route(k, child_p[:, i], p_ev * (dec == i))
        if critic:
            rops = o['router']['n_ops']
            base = ce + k_cpt * (o['n_ops'] + rops)
            o['c_ev'] = base + sum((dec == i) * nodes[k]['c_ev'] for i, k in enumerate(kids))
            o['c_opt'] = base + np.min([nodes[k]['c_opt'] * one for k in kids], 0)
            key = 'c_opt' if hy.get('optimistic', False) else 'c_ev'
            tgt = frozen if frozen is not None else nodes           # stop_gradient(target)
            o['c_cre'] = hy.get('k_cre', 1e-3) * sum(
                (r[:, i] + tgt[k][key]) ** 2 for i, k in enumerate(kids))

    route('', one, one)
    tot = 0.0
    for pth in order:
        o = nodes[pth]
        rmod = o['router']['c_mod'] if o['router'] else 0.0
        rops = o['router']['n_ops'] if o['router'] else 0
        sg_p = frozen[pth]['p_tr'] if frozen is not None else o['p_tr']   # stop_gradient(p_tr)
        if critic:
            tot += np.mean(sg_p * (o['c_err'] + o['c_cre'] + o['c_mod'] + rmod))
        else:
            tot += np.mean(o['p_tr'] * (o['c_err'] + k_cpt * (o['n_ops'] + rops)))
            tot += np.mean(sg_p * (o['c_mod'] + rmod))
            if len(o['rec']['sinks']) > 1:
                tot += np.mean(sg_p * hy.get('k_dec', 0.01) * (o['router']['x'] ** 2).sum(1))
    return dict(c_tot=tot, nodes=nodes, order=order)
