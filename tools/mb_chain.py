"""Dependent-launch latency: a chain of N identical small kernels captured in one CUDA graph (tuning aid)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from lib import _cabi
from lib.engine import _BN_FUSE, _host_struct
from util import Geo
L = _cabi.lib()
vp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
B = int(os.environ.get('B', 128))
N_CHAIN = 40


def chain(name, fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sp = ctypes.c_void_p(s.cuda_stream)
        for _ in range(3):
            fn(sp)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N_CHAIN):
                fn(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    print('%-46s %6.2f us per dependent launch' % (name, 1e3 * float(np.median(ts)) / N_CHAIN), flush=True)


def conv(H, K0, K1, N, stats):
    geo = Geo(B, H, H)
    A0 = torch.randn((K0 // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
    A1 = torch.randn((K1 // 8, geo.P, 8), device='cuda').to(torch.bfloat16) if K1 else None
    Wp = torch.randn((9, (K0 + K1) // 8, N, 8), device='cuda').to(torch.bfloat16)
    out = torch.zeros((N // 8, geo.P, 8), dtype=torch.bfloat16, device='cuda')
    bs = torch.zeros(N, device='cuda')
    acc = torch.zeros(2 * N + 1, dtype=torch.float64, device='cuda')
    gm, bt, ma, va = (torch.ones(N, device='cuda') for _ in range(4))
    ss, mr = torch.zeros(2 * N, device='cuda'), torch.ones(2 * N, device='cuda')
    f = _host_struct(_BN_FUSE, acc=vp(acc), gamma=vp(gm), beta=vp(bt), m_avg=vp(ma), v_avg=vp(va), ss=vp(ss), mr=vp(mr),
                     count=float(B * H * H), d=0.9, eps=1e-6)
    keep.append((A0, A1, Wp, out, bs, acc, gm, bt, ma, va, ss, mr, f))
    if stats:
        return lambda sp: L.conv_bn_stats(vp(A0), K0, vp(A1), K1, vp(Wp), vp(bs), vp(out), N, B, H, H, geo.G, geo.P,
                                          ctypes.c_void_p(f.ctypes.data), 1, 1, sp)
    return lambda sp: L.stencil_gemm(vp(A0), K0, vp(A1), K1, vp(Wp), 9, vp(bs), vp(out), N, 0, None, 0, 0,
                                     B, H, H, geo.G, geo.P, None, 0, None, 1, 1, 1, sp)


def bnf(H, C):
    geo, gp = Geo(B, H, H), Geo(B, H // 2, H // 2)
    lin = torch.randn((C // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
    act = torch.zeros_like(lin)
    pooled = torch.zeros((C // 8, gp.P, 8), dtype=torch.bfloat16, device='cuda')
    ss = torch.randn(2 * C, device='cuda')
    keep.append((lin, act, pooled, ss))
    return lambda sp: L.bn_relu_pool_fwd(vp(lin), C, B, H, H, geo.G, geo.P, vp(ss), vp(act), vp(pooled), gp.P, None, 0, 1, sp)


keep = []
x = torch.zeros(1024, device='cuda')
hyp = torch.zeros(8, device='cuda')
print('B =', B, 'PDL', os.environ.get('MPNN_PDL', '1'))
chain('node_moments (tiny plain kernel)', lambda sp: L.node_moments(vp(x), 1, 1024, vp(hyp), sp))
chain('conv H4  K16+16 N16 + BN stats', conv(4, 16, 16, 16, True))
chain('conv H4  K16+16 N16 plain', conv(4, 16, 16, 16, False))
chain('conv H32 K16 N16 + BN stats', conv(32, 16, 0, 16, True))
chain('conv H8  K64 N64 + BN stats', conv(8, 64, 0, 64, True))
chain('conv H4  K128 N128 + BN stats', conv(4, 128, 0, 128, True))
chain('bn_relu_pool_fwd H32 C16', bnf(32, 16))
chain('bn_relu_pool_fwd H8 C64', bnf(8, 64))
