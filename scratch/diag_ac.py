import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'multipath-nn_b200'), os.path.join(ROOT,'tests')]
import numpy as np, torch
from util import *
from oracle.torch_ref import OracleNet
kind=sys.argv[1] if len(sys.argv)>1 else 'ac'; prec=sys.argv[2] if len(sys.argv)>2 else 'fp32'
net=randomize_routers(tiny_net(kind,k_cpt=4e-9)).configure(precision=prec)
rec=record_of(net); B=24
x0,y=batch(B,seed=3)
o=OracleNet(rec,torch.float64)
for _,_,_,t in o.trainable: t.grad=None
out=o.forward(x0,y,'tr',tau=0.7)
for p in out.order:
    nd=out.nodes[p]
    if nd.router is not None: nd.router.x.retain_grad()
    nd.p_tr.retain_grad() if nd.p_tr.requires_grad else None
out.c_tot.backward()
eng=net._get_engine()
f={net.x0:x0,net.y:y,net.τ:0.7}
eng.train_step(f,update=False); torch.cuda.synchronize()
plan=eng._plan(B,True,True)
paths=node_paths(net)
for nd in eng.switches:
    path=paths[nd.idx][0]
    ref=out.nodes[path].router.x.grad.numpy()
    got=plan.rtr[nd.idx].dR.cpu().numpy()
    print('dR',path,rel_err(got,ref), np.abs(ref).max(), np.abs(got).max())
g=eng.grads_numpy(with_l2=True)
for p,(path,role,key,t) in zip(eng.tparams,o.trainable):
    ref=t.grad.numpy() if t.grad is not None else np.zeros(p.shape)
    n=np.linalg.norm(ref)
    print('%-8s %-6s %-8s ref %.3e err %.3e'%(path,role,key,n,np.linalg.norm(g[p]-ref)/max(n,1e-12)))
