"""Micro-benchmark of the stencil kernels through the C ABI (tuning aid, not a bench line)."""
import ctypes, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from lib import _cabi
from util import Geo

L = _cabi.lib()
vp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
B = int(os.environ.get('B', 2048))
flush = torch.zeros(256 << 20, dtype=torch.uint8, device='cuda')
flush_sink = torch.zeros((), dtype=torch.int64, device='cuda')


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush_sink.copy_(flush.view(torch.int32)[::1].sum(dtype=torch.int64))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3


def conv_case(H, K0, K1, N, stats=True, bias=True, n1=0):
    geo = Geo(B, H, H)
    A0 = torch.randn((K0 // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
    A1 = torch.randn((K1 // 8, geo.P, 8), device='cuda').to(torch.bfloat16) if K1 else None
    Wp = torch.randn((9, (K0 + K1) // 8, N + n1, 8), device='cuda').to(torch.bfloat16)
    out = torch.zeros((N // 8, geo.P, 8), dtype=torch.bfloat16, device='cuda')
    out1 = torch.zeros((n1 // 8, geo.P, 8), dtype=torch.bfloat16, device='cuda') if n1 else None
    bs = torch.zeros(N + n1, device='cuda')
    st = torch.zeros(592 * 2 * (N + n1), device='cuda')
    cnt = ctypes.c_int(0)

    from lib.engine import _BN_FUSE, _BN_BWD_EPI, _host_struct
    acc = torch.zeros(2 * N + 1, dtype=torch.float64, device='cuda')
    gm, bt, ma, va = (torch.ones(N, device='cuda') for _ in range(4))
    ss, mr = torch.zeros(2 * N, device='cuda'), torch.ones(2 * N, device='cuda')
    f = _host_struct(_BN_FUSE, acc=vp(acc), gamma=vp(gm), beta=vp(bt), m_avg=vp(ma), v_avg=vp(va), ss=vp(ss), mr=vp(mr),
                     count=float(B * H * H), d=0.9, eps=1e-6)
    sums, dg, db = torch.zeros(2 * N, device='cuda'), torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda')
    lin = torch.randn((max(N, 8) // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
    epi = _host_struct(_BN_BWD_EPI, lin=vp(lin), ss=vp(ss), mr=vp(mr), acc=vp(acc), sums=vp(sums), dgamma=vp(dg), dbeta=vp(db))

    def run():
        if stats == 2:        # data gradient with the fused BN-backward sums on out0
            L.conv_dgrad_bn_reduce(vp(A0), K0, vp(Wp), vp(out), N, vp(out1), n1, B, H, H, geo.G, geo.P,
                                   ctypes.c_void_p(epi.ctypes.data), 1, 1, None)
        elif stats and not n1:
            L.conv_bn_stats(vp(A0), K0, vp(A1), K1, vp(Wp), vp(bs) if bias else None, vp(out), N, B, H, H, geo.G, geo.P,
                            ctypes.c_void_p(f.ctypes.data), 1, 1, None)
        else:
            L.stencil_gemm(vp(A0), K0, vp(A1), K1, vp(Wp), 9, vp(bs) if bias else None, vp(out), N, 0, vp(out1), n1, 0,
                           B, H, H, geo.G, geo.P, None, 592, ctypes.byref(cnt), 1, 1, 1, None)
    us = timeit(run)
    by = B * H * H * (K0 + K1 + N + n1 + (N if stats == 2 else 0)) * 2
    fl = 2.0 * B * H * H * 9 * (K0 + K1) * (N + n1)
    print('gemm  H%-2d K%d+%d N%d+%d stats=%d bias=%d : %7.1f us  %6.0f GB/s %6.1f TFLOP/s' % (
        H, K0, K1, N, n1, stats, bias, us, by / us / 1e3, fl / us / 1e6), flush=True)


def wgrad_case(H, K0, K1, N):
    geo = Geo(B, H, H)
    A0 = torch.randn((K0 // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
    A1 = torch.randn((K1 // 8, geo.P, 8), device='cuda').to(torch.bfloat16) if K1 else None
    G = torch.randn((N // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
    dW0 = torch.zeros((9, K0, N), device='cuda'); dW1 = torch.zeros((9, K1, N), device='cuda') if K1 else None

    def run():
        L.stencil_wgrad(vp(A0), K0, K0, vp(dW0), vp(A1), K1, K1, vp(dW1), vp(G), N, N, None, 9,
                        B, H, H, geo.G, geo.P, 1, 1, None)
    us = timeit(run)
    by = B * H * H * (K0 + K1 + N) * 2
    fl = 2.0 * B * H * H * 9 * (K0 + K1) * N
    print('wgrad H%-2d K%d+%d N%d : %7.1f us  %6.0f GB/s %6.1f TFLOP/s' % (H, K0, K1, N, us, by / us / 1e3, fl / us / 1e6),
          flush=True)


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    print('B =', B, 'per_sm', os.environ.get('MPNN_TUNE_PER_SM'), 'nstage', os.environ.get('MPNN_TUNE_NSTAGE'))
    if what in ('all', 'gemm'):
        conv_case(32, 16, 0, 16)
        conv_case(32, 16, 0, 16, stats=False)
        conv_case(32, 16, 0, 16, stats=False, bias=False)
        conv_case(16, 16, 16, 16)
        conv_case(16, 32, 0, 32)
        conv_case(8, 64, 0, 64)
        conv_case(4, 128, 0, 128)
        conv_case(4, 16, 16, 16)
        conv_case(32, 16, 0, 0, stats=False, bias=False, n1=16) if False else None
    if what == 'r2':
        conv_case(32, 16, 0, 16)                                  # stage-0/1 forward, fused BN moments
        conv_case(32, 16, 0, 16, stats=False, bias=False)         # plain data gradient
        conv_case(32, 16, 0, 16, stats=2, bias=False)             # data gradient + BN-backward sums
        conv_case(16, 16, 16, 16)
        conv_case(16, 16, 0, 16, stats=2, bias=False, n1=16)
        conv_case(16, 32, 0, 32)
        conv_case(16, 32, 0, 32, stats=2, bias=False)
        conv_case(8, 32, 32, 32)
        conv_case(8, 64, 0, 64)
        conv_case(8, 64, 0, 64, stats=2, bias=False)
        conv_case(4, 64, 64, 64)
        conv_case(4, 128, 0, 128)
        conv_case(4, 16, 16, 16)
    if what == 'h32fwd':
        conv_case(32, 16, 0, 16)
    if what == 'h32dgrad':
        conv_case(32, 16, 0, 16, stats=2, bias=False)
    if what == 'h32wgrad':
        wgrad_case(32, 16, 0, 16)
    if what == 'h8wgrad':
        wgrad_case(8, 64, 0, 64)
    if what == 'h4':                                              # coarse-scale layers at the reference's batch (B=128)
        conv_case(4, 64, 64, 64)
        conv_case(4, 64, 64, 64, stats=False)
        conv_case(4, 128, 0, 128)
        conv_case(4, 128, 0, 64, stats=False, bias=False, n1=0)
        wgrad_case(4, 64, 64, 64)
    if what == 'gemm1':
        conv_case(32, 16, 0, 16)
        conv_case(16, 32, 0, 32)
        conv_case(8, 64, 0, 64)
    if what == 'nsweep':
        for n in (16, 32, 48, 64, 96):
            conv_case(32, 16, 0, n, stats=False)
    if what in ('all', 'wgrad'):
        wgrad_case(32, 16, 0, 16)
        wgrad_case(16, 16, 16, 16)
        wgrad_case(16, 32, 0, 32)
        wgrad_case(8, 64, 0, 64)
        wgrad_case(4, 128, 0, 128)
        wgrad_case(4, 16, 16, 16)
