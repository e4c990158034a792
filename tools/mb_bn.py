"""Micro-benchmark of the BatchNorm/ReLU/pool kernels through the C ABI (tuning aid)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from lib import _cabi
from util import Geo
from mb_conv import timeit, vp, L, B


def case(H, C, pool=True, dt=1):
    geo, gp = Geo(B, H, H), Geo(B, H // 2, H // 2)
    td = torch.bfloat16 if dt else torch.float32
    es = 2 if dt else 4
    lin = torch.randn((C // 8, geo.P, 8), device='cuda').to(td)
    act = torch.zeros_like(lin); dact = torch.randn_like(lin.float()).to(td); dlin = torch.zeros_like(lin)
    pooled = torch.zeros((C // 8, gp.P, 8), dtype=td, device='cuda') if pool else None
    dpooled = torch.randn((C // 8, gp.P, 8), device='cuda').to(td) if pool else None
    ss = torch.randn(2 * C, device='cuda'); mr = torch.rand(2 * C, device='cuda') + 0.5
    sums = torch.randn(2 * C, device='cuda'); parts = torch.zeros(592 * 2 * C, device='cuda')
    dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda'); dbias = torch.zeros(C, device='cuda')
    cnt = ctypes.c_int(0)
    n = B * H * H * C * es
    t = timeit(lambda: L.bn_relu_pool_fwd(vp(lin), C, B, H, H, geo.G, geo.P, vp(ss), vp(act), vp(pooled), gp.P if pool else 0,
                                          None, 0, dt, None))
    print('bn_fwd     H%-2d C%-3d pool=%d: %6.1f us %6.0f GB/s' % (H, C, pool, t, n * (2 + 0.25 * pool) / t / 1e3), flush=True)
    t = timeit(lambda: L.bn_bwd_reduce(vp(lin), vp(dact), None, 0, vp(ss), vp(mr), C, B, H, H, geo.G, geo.P,
                                       vp(parts), 592, ctypes.byref(cnt), dt, None))
    print('bn_bwd_red H%-2d C%-3d       : %6.1f us %6.0f GB/s' % (H, C, t, n * 2 / t / 1e3), flush=True)
    t = timeit(lambda: L.bn_bwd_finalize(vp(parts), cnt.value, C, vp(mr), vp(sums), vp(dg), vp(db), None))
    print('bn_bwd_fin H%-2d C%-3d       : %6.1f us' % (H, C, t), flush=True)
    t = timeit(lambda: L.bn_relu_pool_bwd(vp(lin), vp(dact), None, 0, vp(dpooled), gp.P if pool else 0, vp(ss), vp(mr), vp(sums),
                                          float(B * H * H), C, B, H, H, geo.G, geo.P, vp(dlin), vp(dbias), dt, None))
    print('bn_bwd     H%-2d C%-3d pool=%d: %6.1f us %6.0f GB/s' % (H, C, pool, t, n * (3 + 0.25 * pool) / t / 1e3), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'h32':
        case(32, 16)
    else:
        case(32, 16); case(16, 32); case(16, 16); case(8, 64); case(4, 128, pool=False)
