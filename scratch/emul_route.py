import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'multipath-nn_b200'), os.path.join(ROOT,'tests')]
import numpy as np, torch
from util import *
from oracle.torch_ref import OracleNet
from lib.net_types import n_leaves
net=randomize_routers(tiny_net('ac',k_cpt=4e-9)); rec=record_of(net); B=24
x0,y=batch(B,seed=3)
o=OracleNet(rec,torch.float64)
out=o.forward(x0,y,'tr',tau=0.7)
for p in out.order:
    nd=out.nodes[p]
    if nd.router is not None: nd.router.x.retain_grad()
out.c_tot.backward()
paths=[p for p,_ in node_paths(net)]; layers=[l for _,l in node_paths(net)]
idx={p:i for i,p in enumerate(paths)}
n=len(paths); tau=0.7; eps=1e-6; kc=4e-9; kdec=0.01
parent=[-1]*n; sink=[0]*n; kids=[[] for _ in range(n)]
for p in paths:
    if p=='' : continue
    par=p.rsplit('/',1)[0] if '/' in p else ''
    parent[idx[p]]=idx[par]; sink[idx[p]]=int(p.rsplit('/',1)[-1]); kids[idx[par]].append(idx[p])
root_leaves=n_leaves(net.root)
floor=[n_leaves(l)/root_leaves for l in layers]
ops=[l.n_ops+(l.router.n_ops if l.router is not None else 0) for l in layers]
def sm(r): 
    x=r/tau; e=np.exp(x-x.max(1,keepdims=True)); return e/e.sum(1,keepdims=True)
gp=np.zeros((n,B))
for i,p in enumerate(paths):
    nd=out.nodes[p]
    ce=nd.c_err.detach().numpy() if torch.is_tensor(nd.c_err) else 0.0
    gp[i]=(ce+kc*ops[i])/B
for i in range(n-1,0,-1):
    par=parent[i]; g=gp[i].copy()
    if len(kids[par])>=2:
        r=out.nodes[paths[par]].router.x.detach().numpy(); g=g*sm(r)[:,sink[i]]
    gp[par]+=g
for i,p in enumerate(paths):
    if len(kids[i])<2: continue
    r=out.nodes[p].router.x.detach().numpy(); s=sm(r); pt=out.nodes[p].p_tr.detach().numpy()
    gs=np.stack([gp[c]*(pt-eps*floor[i]) for c in kids[i]],1)
    dot=(s*gs).sum(1,keepdims=True)
    dR=s*(gs-dot)/tau+pt[:,None]*kdec*2*r/B
    ref=out.nodes[p].router.x.grad.numpy()
    print(p, rel_err(dR,ref))
