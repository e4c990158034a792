"""Architecture, hyperparameters and net constructors -- the values and
builder names of /root/reference/scripts/arch_and_hypers.py:12-139.

Difference: `dr_tree`'s helper takes `y_shape` explicitly (in the reference it
resolves to an undefined global and every *-tree experiment raises NameError,
SURVEY F5).
"""
from lib.layer_types import (
    BatchNorm, Chain, CrossEntropyError, LinTrans, MultiscaleBatchNorm,
    MultiscaleConvMax, MultiscaleLLN, MultiscaleRect, Rect, Select,
    Softmax, ToPyramid)
from lib.net_types import CriticNet, ActorNet, SRNet

# ---- network hyperparameters ------------------------------------------------

conv_supp = 3
router_n_chan = 16

k_cpts = [0.0, 1e-9, 2e-9, 4e-9, 8e-9, 1.6e-8, 3.2e-8, 6.4e-8]
k_l2 = 1e-4
σ_w = 1

arch = [[16] * 4, [16] * 4, [32] * 3, [32] * 3, [64] * 2, [64] * 2, [128], [128]]

# ---- training hyperparameters -----------------------------------------------

n_iter = 80000
t_log = 2500
batch_size = 128


def λ_lrn(t):
    return 0.1 / 2 ** (t / 10000)


def τ_cr(t):
    return 0.1 / 2 ** (t / 20000)


def τ_ds(t):
    return 1 / 2 ** (t / 20000)

# ---- network components -----------------------------------------------------


def _fc(n, σ=σ_w):
    return LinTrans(n_chan=n, k_l2=k_l2, σ_w=σ)


def router(n_sinks):
    if n_sinks < 2:
        return None
    return Chain(name='Router', comps=[
        Select(i=-1), _fc(router_n_chan), BatchNorm(), Rect(),
        _fc(router_n_chan), BatchNorm(), Rect(), _fc(n_sinks, 0)])


def pyr(*sinks):
    return Chain(name='ToPyramid', sinks=sinks, router=router(len(sinks)),
                 comps=[ToPyramid(n_scales=len(arch[0]))])


def rcm(i, *sinks):
    return Chain(name='ReConvMax', sinks=sinks, router=router(len(sinks)), comps=[
        MultiscaleConvMax(n_chan=arch[i], supp=conv_supp, k_l2=k_l2, σ_w=σ_w),
        MultiscaleBatchNorm(), MultiscaleRect()])


def reg(n_chan):
    return Chain(name='LogReg', comps=[
        Select(i=-1), _fc(n_chan), Softmax(), CrossEntropyError()])

# ---- network constructors ---------------------------------------------------


def sr_chain(n_tf):
    def make_net(x0_shape, y_shape):
        node = reg(y_shape[0])
        for i in range(n_tf - 1, -1, -1):
            node = rcm(i, node)
        return SRNet(x0_shape=x0_shape, y_shape=y_shape, root=pyr(node))
    return make_net


def dr_chain(type_, **hypers):
    def make_net(x0_shape, y_shape):
        node = rcm(-1, reg(y_shape[0]))
        for i in range(len(arch) - 2, -1, -1):
            node = rcm(i, reg(y_shape[0]), node)
        return type_(x0_shape=x0_shape, y_shape=y_shape, root=pyr(node), **hypers)
    return make_net


def dr_tree(type_, **hypers):
    def tail(n_cls, first=3):
        node = rcm(7, reg(n_cls))
        for i in range(6, first - 1, -1):
            node = rcm(i, reg(n_cls), node)
        return node

    def make_net(x0_shape, y_shape):
        c = y_shape[0]

        def stage2():
            return rcm(2, reg(c), tail(c), tail(c))

        def stage1():
            return rcm(1, reg(c), stage2(), stage2())
        root = pyr(rcm(0, reg(c), stage1(), stage1()))
        return type_(x0_shape=x0_shape, y_shape=y_shape, root=root, **hypers)
    return make_net


def ac_chain(**hypers):
    return dr_chain(ActorNet, **hypers)


def ac_tree(**hypers):
    return dr_tree(ActorNet, **hypers)


def cr_chain(**hypers):
    return dr_chain(CriticNet, **hypers)


def cr_tree(**hypers):
    return dr_tree(CriticNet, **hypers)
