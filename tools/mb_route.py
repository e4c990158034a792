"""Routing / compaction / gather / scatter / BN / head kernels at B = 4096 for one ncu capture (tuning aid)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
from lib import layer_types, _cabi
import arch_and_hypers as ah
from util import Geo

B = int(os.environ.get('B', 4096))
L = _cabi.lib()
vp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
layer_types.seed(0)
net = ah.ac_chain(k_cpt=4e-9)((32, 32, 3), (10,)).configure(precision='bf16', graphs=os.environ.get('GRAPHS', '1') != '0')
rng = np.random.default_rng(1)
for l in net.layers:
    if l.router is not None:
        w = l.router.comps[-1].params.w
        w.assign((0.5 * rng.standard_normal(w.shape)).astype(np.float32))
x0 = rng.random((B, 32, 32, 3)).astype(np.float32)
y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
for t in range(3):
    net.train.run({net.x0: x0, net.y: y, net.τ: 1.0, net.mode: 'tr', net.λ_lrn: 0.0})
torch.cuda.synchronize()
ev = net.compact_evaluator(B)
ev.reset()
ev.run_batch(x0, y)
torch.cuda.synchronize()
print('visits per node', ev.visits.tolist())
# scatter-add / gather of whole image blocks (stage-1 -> stage-2 hand-over: 16 channels at 16x16)
geo = Geo(B, 16, 16)
src = torch.randn((2, geo.P, 8), device='cuda').to(torch.bfloat16)
dst = torch.zeros_like(src)
idx = torch.from_numpy(rng.permutation(B)[:B // 2].astype(np.int32)).cuda()
cnt = torch.tensor([B // 2], dtype=torch.int32, device='cuda')
for _ in range(2):
    L.gather_images(vp(src), B, geo.P, vp(idx), vp(cnt), vp(dst), B, geo.P, 16, 16, 16, geo.G, 1, None)
    L.scatter_add_images(vp(dst), B, geo.P, vp(idx), vp(cnt), vp(src), B, geo.P, 16, 16, 16, geo.G, 1, None)
p_ev = torch.zeros((16, B), device='cuda'); p_ev[:, ::2] = 1
cidx = torch.zeros((16, B), dtype=torch.int32, device='cuda'); ccnt = torch.zeros(16, dtype=torch.int32, device='cuda')
L.compact_paths(vp(p_ev), 16, B, vp(cidx), vp(ccnt), None)
torch.cuda.synchronize()
print('ok')
