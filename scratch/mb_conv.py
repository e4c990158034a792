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

    def run():
        L.stencil_gemm(vp(A0), K0, vp(A1), K1, vp(Wp), 9, vp(bs) if bias else None, vp(out), N, 0, vp(out1), n1, 0,
                       B, H, H, geo.G, geo.P, vp(st) if stats else None, 592, ctypes.byref(cnt), 1, 1, 1, None)
    us = timeit(run)
    by = B * H * H * (K0 + K1 + N + n1) * 2
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
