"""Per-tile pipeline timeline of CTA 0 of the conv kernel (MPNN_TUNE_DBG=32)."""
import ctypes, os, sys
os.environ['MPNN_TUNE_DBG'] = str(32 | int(os.environ.get('DBG_EXTRA', '0')))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from mb_conv import L, vp, B
from util import Geo

H, K0, N = 32, 16, 16
geo = Geo(B, H, H)
A0 = torch.randn((K0 // 8, geo.P, 8), device='cuda').to(torch.bfloat16)
Wp = torch.randn((9, K0 // 8, N, 8), device='cuda').to(torch.bfloat16)
out = torch.zeros((N // 8, geo.P, 8), dtype=torch.bfloat16, device='cuda')
bs = torch.zeros(N, device='cuda')
st = torch.zeros(592 * 2 * N, dtype=torch.int32, device='cuda')
cnt = ctypes.c_int(0)
for rep in range(3):
    st.zero_()
    L.stencil_gemm(vp(A0), K0, None, 0, vp(Wp), 9, vp(bs), vp(out), N, 0, None, 0, 0,
                   B, H, H, geo.G, geo.P, vp(st), 592, ctypes.byref(cnt), 1, 1, 1, None)
    torch.cuda.synchronize()
t = st.cpu().numpy().astype(np.uint32)[:64 * 8].reshape(64, 8).astype(np.int64)
t0 = t[0, 0]
names = ['load_issue', 'mma_tempty', 'mma_full', 'mma_commit', 'epi_tfull', 'epi_done']
print('tile ' + ' '.join('%11s' % n for n in names) + '   | full-issue  commit-full  tfull-commit  epi')
for i in range(28):
    r = (t[i, :6] - t0) & 0xffffffff
    print('%4d ' % i + ' '.join('%11d' % v for v in r) + '   | %9d %11d %12d %5d' % (
        r[2] - r[0], r[3] - r[2], r[4] - r[3], r[5] - r[4]))
