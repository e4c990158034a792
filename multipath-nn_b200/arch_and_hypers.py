"""The experiments' architecture, hyper-parameters, schedules and net constructors.

Values and public names follow /root/reference/scripts/arch_and_hypers.py:12-139, because the
drivers and the parity tests address them by name (`arch`, `k_cpts`, `λ_lrn`, `τ_ds`, `τ_cr`,
`router / pyr / rcm / reg`, `sr_chain`, `dr_chain`, `dr_tree`, `ac_* / cr_*`).  One difference:
the tree builder takes the class count from `y_shape` (in the reference that name is an
undefined global and every *-tree experiment raises NameError).

Topology: a 4-scale image pyramid feeds eight "ReConvMax" stages; stage i has
`len(arch[i])` scales of `arch[i][0]` channels, so the pyramid narrows from 32x32..4x4 x16 to
4x4 x128.  In the dynamically-routed nets every stage but the last owns a LogReg exit and a
2-way router (exit here / continue); the trees branch 2-ways below stages 0, 1 and 2.
"""
from lib.layer_types import (
    BatchNorm, Chain, CrossEntropyError, LinTrans, MultiscaleBatchNorm,
    MultiscaleConvMax, MultiscaleLLN, MultiscaleRect, Rect, Select,
    Softmax, ToPyramid)
from lib.net_types import ActorNet, CriticNet, SRNet

# ------------------------------------------------------------------ network ---
_STAGE_WIDTH = (16, 16, 32, 32, 64, 64, 128, 128)
_STAGE_SCALES = (4, 4, 3, 3, 2, 2, 1, 1)
arch = [[width] * scales for width, scales in zip(_STAGE_WIDTH, _STAGE_SCALES)]

conv_supp = 3            # 3x3 kernels
router_n_chan = 16       # width of the two hidden router layers
k_l2 = 1e-4              # weight decay of every conv / FC weight
σ_w = 1                  # initialisation scale
k_cpts = [0.0] + [1e-9 * 2 ** i for i in range(7)]      # cost-of-computation sweep: 0, 1e-9 ... 6.4e-8

# ----------------------------------------------------------------- training ---
n_iter, t_log, batch_size = 80000, 2500, 128


def _halving(start, period):
    """t -> start * 2^(-t / period)"""
    return lambda t: start / 2 ** (t / period)


λ_lrn = _halving(0.1, 10000)     # learning rate
τ_ds = _halving(1.0, 20000)      # actor routing temperature
τ_cr = _halving(0.1, 20000)      # critic routing temperature


# --------------------------------------------------------------- components ---
def _dense(width, scale=σ_w):
    return LinTrans(n_chan=width, k_l2=k_l2, σ_w=scale)


def router(n_sinks):
    """FC16-BN-ReLU-FC16-BN-ReLU-FC(n_sinks, zero-initialised) on the coarsest scale; only switches have one"""
    if n_sinks < 2:
        return None
    hidden = []
    for _ in range(2):
        hidden += [_dense(router_n_chan), BatchNorm(), Rect()]
    return Chain(name='Router', comps=[Select(i=-1)] + hidden + [_dense(n_sinks, 0)])


def _node(name, comps, sinks):
    return Chain(name=name, comps=comps, sinks=sinks, router=router(len(sinks)))


def pyr(*sinks):
    return _node('ToPyramid', [ToPyramid(n_scales=_STAGE_SCALES[0])], sinks)


def rcm(i, *sinks):
    conv = MultiscaleConvMax(n_chan=arch[i], supp=conv_supp, k_l2=k_l2, σ_w=σ_w)
    return _node('ReConvMax', [conv, MultiscaleBatchNorm(), MultiscaleRect()], sinks)


def reg(n_chan):
    return Chain(name='LogReg', comps=[Select(i=-1), _dense(n_chan), Softmax(), CrossEntropyError()])


# ------------------------------------------------------------- constructors ---
def _stack(stages, below):
    """stages[0] -> stages[1] -> ... -> below, every stage built by `make(i, child)`"""
    node = below
    for make in reversed(stages):
        node = make(node)
    return node


def sr_chain(n_tf):
    """statically-routed: the first n_tf stages, one classifier at the end"""
    def make_net(x0_shape, y_shape):
        stages = [(lambda child, i=i: rcm(i, child)) for i in range(n_tf)]
        return SRNet(x0_shape=x0_shape, y_shape=y_shape, root=pyr(_stack(stages, reg(y_shape[0]))))
    return make_net


def dr_chain(type_, **hypers):
    """dynamically-routed chain: every stage but the last chooses between its own classifier and the next stage"""
    def make_net(x0_shape, y_shape):
        n_cls = y_shape[0]
        stages = [(lambda child, i=i: rcm(i, reg(n_cls), child)) for i in range(len(arch) - 1)]
        return type_(x0_shape=x0_shape, y_shape=y_shape, root=pyr(_stack(stages, rcm(-1, reg(n_cls)))), **hypers)
    return make_net


def dr_tree(type_, **hypers):
    """dynamically-routed tree: stages 0, 1, 2 choose between their classifier and TWO copies of the rest"""
    def make_net(x0_shape, y_shape):
        n_cls = y_shape[0]

        def chain_from(first):                       # stages first..7 as a chain of 2-way switches
            stages = [(lambda child, i=i: rcm(i, reg(n_cls), child)) for i in range(first, 7)]
            return _stack(stages, rcm(7, reg(n_cls)))

        def branch(i):                               # stage i with two identical subtrees below it
            below = (lambda: branch(i + 1)) if i < 2 else (lambda: chain_from(3))
            return rcm(i, reg(n_cls), below(), below())
        return type_(x0_shape=x0_shape, y_shape=y_shape, root=pyr(branch(0)), **hypers)
    return make_net


def ac_chain(**hypers):
    return dr_chain(ActorNet, **hypers)


def ac_tree(**hypers):
    return dr_tree(ActorNet, **hypers)


def cr_chain(**hypers):
    return dr_chain(CriticNet, **hypers)


def cr_tree(**hypers):
    return dr_tree(CriticNet, **hypers)
