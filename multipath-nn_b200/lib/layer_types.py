"""Layer types of the B200 build -- same names, hypers, `params` keys and
`link(x, y, mode)` protocol as the reference's `lib/layer_types.py`
(/root/reference/scripts/lib/layer_types.py), but nothing here builds a
TensorFlow graph: `link` performs shape inference, creates the parameters
with the reference's initialisation formulas and records `n_ops`; the
arithmetic is executed by `lib.engine` through libmpnn_sm100 (CUDA, sm_100a).

`x` handed to `link` is a `Sym` (a symbolic tensor: shape without the batch
dimension + its producer) or a list of them for multiscale pyramids.
"""
from types import SimpleNamespace as Ns

import numpy as np

__all__ = [
    'Sym', 'Param', 'seed', 'Layer', 'NoOp', 'LinTrans', 'Conv', 'Rect', 'Softmax', 'MaxPool',
    'GlobalMaxPool', 'ToPyramid', 'MultiscaleLLN', 'MultiscaleConvMax', 'MultiscaleRect', 'Select',
    'Dropout', 'BatchNorm', 'MultiscaleBatchNorm', 'SquaredError', 'CrossEntropyError',
    'SuperclassCrossEntropyError', 'ActivityError', 'Chain']

# ---- Symbols and Parameters ------------------------------------------------

_rng = np.random.default_rng(0)


def seed(s):
    """Seed the parameter initialiser (the reference draws from an unseeded
    tf.random_normal; a seed is needed for reproducible parity runs)."""
    global _rng
    _rng = np.random.default_rng(s)


def _normal(shape, scale):
    return (scale * _rng.standard_normal(shape)).astype(np.float32)


class Sym:
    """Symbolic tensor: `shape` excludes the batch dimension."""

    def __init__(self, shape, owner=None, name='x'):
        self.shape = tuple(int(s) for s in shape)
        self.owner = owner
        self.name = name

    def __repr__(self):
        return 'Sym(%s.%s %s)' % (getattr(self.owner, 'name', None), self.name, self.shape)


class Param:
    """Replaces tf.Variable: `.eval()` returns the current value as numpy
    (fetched from the device once an engine owns the parameter)."""

    def __init__(self, value, trainable=True):
        self.value = np.ascontiguousarray(value, dtype=np.float32)
        self.trainable = trainable
        self._bind = None          # (engine, slot) once uploaded

    @property
    def shape(self):
        return self.value.shape

    def eval(self):
        if self._bind is not None:
            self._bind[0].fetch_param(self)
        return self.value.copy()

    def assign(self, v):
        v = np.asarray(v, dtype=np.float32)
        if v.shape != self.value.shape:
            raise ValueError('assign: shape %s != %s' % (v.shape, self.value.shape))
        self.value = np.ascontiguousarray(v)
        if self._bind is not None:
            self._bind[0].push_param(self)


def _shape(x):
    return [s.shape for s in x] if isinstance(x, list) else x.shape

# ---- Core Layer Class ------------------------------------------------------


class Layer:
    default_hypers = Ns()

    def __init__(self, **options):
        self.name = options.pop('name', type(self).__name__)
        self.router = options.pop('router', None)
        self.sinks = list(options.pop('sinks', []))
        self.comps = list(options.pop('comps', []))
        self.hypers = Ns(**{**vars(type(self).default_hypers), **options})
        self.params = Ns()

    def link(self, x, y, mode):
        self.x_in = x
        self.x = x
        self.c_err = 0.0
        self.c_mod = 0.0
        self.n_ops = 0

    def _param(self, key, make, trainable=True):
        """Create `params.<key>` unless a value is already there with the right
        shape (relinking a decoded net must not redraw the weights)."""
        old = getattr(self.params, key, None)
        new = make()
        if old is not None and old.shape == new.shape:
            return old
        p = Param(new, trainable)
        setattr(self.params, key, p)
        return p

# ---- The No-Op Layer -------------------------------------------------------


class NoOp(Layer):
    pass

# ---- Transformation Layers -------------------------------------------------


class LinTrans(Layer):
    default_hypers = Ns(n_chan=1, k_l2=0, σ_w=1, res=False)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        ϕ = self.hypers
        n_in = int(np.prod(x.shape))
        w_scale = ϕ.σ_w / np.sqrt(n_in)
        w_eq = np.eye(n_in, ϕ.n_chan, dtype=np.float32) if ϕ.res else 0
        self._param('w', lambda: w_eq + _normal((n_in, ϕ.n_chan), w_scale))
        self._param('b', lambda: np.zeros(ϕ.n_chan, np.float32))
        self.x = Sym((ϕ.n_chan,), self)
        self.n_ops = n_in * ϕ.n_chan


class Conv(Layer):
    default_hypers = Ns(n_chan=1, supp=1, k_l2=0, σ_w=1, res=False)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        ϕ = self.hypers
        h, w, n_in = x.shape
        w_scale = ϕ.σ_w / ϕ.supp / np.sqrt(n_in)
        mid = (np.arange(ϕ.supp) == ϕ.supp // 2)
        w_eq = (np.float32(mid[:, None, None, None] * mid[:, None, None] * np.eye(n_in, ϕ.n_chan))
                if ϕ.res else 0)
        self._param('w', lambda: w_eq + _normal((ϕ.supp, ϕ.supp, n_in, ϕ.n_chan), w_scale))
        self._param('b', lambda: np.zeros(ϕ.n_chan, np.float32))
        self.x = Sym((h, w, ϕ.n_chan), self)
        self.n_ops = h * w * ϕ.supp ** 2 * n_in * ϕ.n_chan


class Rect(Layer):
    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = Sym(x.shape, self)


class Softmax(Layer):
    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = Sym(x.shape, self)


class MaxPool(Layer):
    default_hypers = Ns(stride=1, supp=1)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        ϕ = self.hypers
        # the reference hands (strides, k_shape) to tf.nn.max_pool(value, ksize,
        # strides): the step between windows is `supp`
        h, w, c = x.shape
        self.x = Sym((-(-h // ϕ.supp), -(-w // ϕ.supp), c), self)


class GlobalMaxPool(Layer):
    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = Sym(x.shape[-1:], self)

# ---- Multiscale Transformation Layers --------------------------------------


class ToPyramid(Layer):
    default_hypers = Ns(n_scales=1)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        h, w, c = x.shape
        self.x = [Sym((h // 2 ** i, w // 2 ** i, c), self, 'x[%i]' % i)
                  for i in range(self.hypers.n_scales)]


class MultiscaleLLN(Layer):
    default_hypers = Ns(shape0=(1, 1), σ=3, ϵ=1e-3)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = [Sym(s.shape, self, 'x[%i]' % i) for i, s in enumerate(x)]


class MultiscaleConvMax(Layer):
    default_hypers = Ns(n_chan=[], supp=1, k_l2=0, σ_w=1)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        ϕ = self.hypers
        n = len(ϕ.n_chan)
        x_in = x[len(x) - n:]
        self.n_ops = 0
        outs = []
        for k, s in enumerate(x_in):
            h, w, c_in = s.shape
            kh, kw = min(ϕ.supp, h), min(ϕ.supp, w)
            self._param('w_horz_%i' % k, lambda: _normal(
                (kh, kw, c_in, ϕ.n_chan[k]), ϕ.σ_w / ϕ.supp / np.sqrt(c_in)))
            per_px = kh * kw * c_in * ϕ.n_chan[k]
            if k > 0:
                per_px += ϕ.supp * ϕ.supp * ϕ.n_chan[k - 1] * ϕ.n_chan[k]
            self.n_ops += h * w * per_px
            outs.append(Sym((h, w, ϕ.n_chan[k]), self, 'x[%i]' % k))
        for k in range(n - 1):
            self._param('w_vert_%i' % k, lambda: _normal(
                (ϕ.supp, ϕ.supp, ϕ.n_chan[k], ϕ.n_chan[k + 1]),
                ϕ.σ_w / ϕ.supp / np.sqrt(ϕ.n_chan[k])))
        for k in range(n):
            self._param('b_%i' % k, lambda: np.zeros(ϕ.n_chan[k], np.float32))
        # keep the reference's key order: w_horz_*, w_vert_*, b_*
        θ = vars(self.params)
        order = (['w_horz_%i' % k for k in range(n)] + ['w_vert_%i' % k for k in range(n - 1)]
                 + ['b_%i' % k for k in range(n)])
        self.params = Ns(**{k: θ[k] for k in order})
        self.x = outs


class MultiscaleRect(Layer):
    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = [Sym(s.shape, self, 'x[%i]' % i) for i, s in enumerate(x)]


class Select(Layer):
    default_hypers = Ns(i=0)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = x[self.hypers.i]

# ---- Regularization Layers -------------------------------------------------


class Dropout(Layer):
    default_hypers = Ns(λ=1)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.x = Sym(x.shape, self)


class BatchNorm(Layer):
    default_hypers = Ns(d=0.9, ϵ=1e-6)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        n_chan = x.shape[-1]
        self._param('γ', lambda: np.ones(n_chan, np.float32))
        self._param('β', lambda: np.zeros(n_chan, np.float32))
        self._param('m_avg', lambda: np.zeros(n_chan, np.float32), trainable=False)
        self._param('v_avg', lambda: np.ones(n_chan, np.float32), trainable=False)
        self.x = Sym(x.shape, self)


class MultiscaleBatchNorm(Layer):
    default_hypers = Ns(d=0.9, ϵ=1e-6)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        # the reference replaces comps at every link (layer_types.py:246);
        # decoded comps (same count) are kept so their loaded params survive
        if len(self.comps) != len(x) or not all(isinstance(c, BatchNorm) for c in self.comps):
            self.comps = [BatchNorm() for _ in x]
        for ℓ, x_i in zip(self.comps, x):
            ℓ.link(x_i, y, mode)
        self.x = [ℓ.x for ℓ in self.comps]

# ---- Error Layers ----------------------------------------------------------


class SquaredError(Layer):
    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.c_err = Sym((), self, 'c_err')
        self.δ_cor = Sym((), self, 'δ_cor')


class CrossEntropyError(Layer):
    default_hypers = Ns(ϵ=1e-6)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.c_err = Sym((), self, 'c_err')
        self.δ_cor = Sym((), self, 'δ_cor')


class SuperclassCrossEntropyError(Layer):
    default_hypers = Ns(w_cls=None, ϵ=1e-6)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.c_err = Sym((), self, 'c_err')
        self.δ_cor = Sym((), self, 'δ_cor')


class ActivityError(Layer):
    default_hypers = Ns(α=0.0)

    def link(self, x, y, mode):
        super().link(x, y, mode)
        self.c_mod = Sym((), self, 'c_mod')

# ---- Compound Layers -------------------------------------------------------


class Chain(Layer):
    def link(self, x, y, mode):
        super().link(x, y, mode)
        for ℓ in self.comps:
            ℓ.link(x, y, mode)
            x = ℓ.x
        self.x = x
        self.n_ops = sum(ℓ.n_ops for ℓ in self.comps)
        errs = [ℓ for ℓ in self.comps if isinstance(ℓ.c_err, Sym)]
        if errs:
            self.c_err = Sym((), self, 'c_err')
        if len(self.comps) > 0 and hasattr(self.comps[-1], 'δ_cor'):
            self.δ_cor = Sym((), self, 'δ_cor')
