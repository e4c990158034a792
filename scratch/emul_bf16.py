"""How much gradient error does bf16 STORAGE (activations, gradients, packed weights) cause by itself?"""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'multipath-nn_b200'), os.path.join(ROOT,'tests')]
import numpy as np, torch
from util import *
from oracle import torch_ref
from oracle.torch_ref import OracleNet, tf_conv2d_same, tf_max_pool_same
class Rnd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.to(torch.bfloat16).to(x.dtype)
    @staticmethod
    def backward(ctx, g): return g.to(torch.bfloat16).to(g.dtype)
class RndF(torch.autograd.Function):   # round forward only (weights)
    @staticmethod
    def forward(ctx, x): return x.to(torch.bfloat16).to(x.dtype)
    @staticmethod
    def backward(ctx, g): return g
MODE=sys.argv[1] if len(sys.argv)>1 else 'all'
class BfNet(OracleNet):
    def _link(self, rec, x, y, mode):
        kind=rec['type']; p=self.params[id(rec)]; hy=rec['hypers']
        if kind=='MultiscaleConvMax':
            n=len(hy['n_chan']); xin=x[len(x)-n:]; out=[]
            nd=torch_ref._Node(x=None,c_err=0.0,c_mod=0.0,n_ops=0,delta_cor=None,comps=[])
            for k in range(n):
                wh=RndF.apply(p['w_horz_%i'%k]); o=p['b_%i'%k]+tf_conv2d_same(xin[k],wh)
                if k>0:
                    wv=RndF.apply(p['w_vert_%i'%(k-1)])
                    pooled=tf_max_pool_same(out[k-1],2,2)
                    if MODE in('all',): pooled=Rnd.apply(pooled)
                    o=o+tf_conv2d_same(pooled,wv)
                if MODE in ('all','lin'): o=Rnd.apply(o)
                out.append(o)
            nd.x=out; return nd
        nd=super()._link(rec,x,y,mode)
        if kind=='MultiscaleRect' and MODE in('all','act'): nd.x=[Rnd.apply(v) for v in nd.x]
        if kind=='ToPyramid': nd.x=[RndF.apply(v) for v in nd.x]
        return nd
full = len(sys.argv)>2
if full:
    sys.path.insert(0, os.path.join(ROOT,'multipath-nn_b200'))
    import arch_and_hypers as ah
    from lib import layer_types; layer_types.seed(0)
    net=ah.ac_chain(k_cpt=4e-9)((32,32,3),(10,)); x0,y=batch(16,(32,32,3),seed=3)
else:
    net=tiny_net('ac',k_cpt=4e-9); x0,y=batch(24,seed=3)
randomize_routers(net); rec=record_of(net)
o=OracleNet(rec,torch.float64); _,g=o.grads(x0,y,tau=0.7)
b=BfNet(rec,torch.float64); _,gb=b.grads(x0,y,tau=0.7)
for (k1,v1),(k2,v2) in zip(g.items(),gb.items()):
    n=v1.norm().item()
    if n>1e-9 and k1[2].startswith('w_horz'): print(k1[:3],'%.4f'%((v1-v2).norm().item()/n))
