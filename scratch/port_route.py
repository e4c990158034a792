import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'multipath-nn_b200'), os.path.join(ROOT,'tests')]
import numpy as np, torch
from util import tiny_net, rel_err
from lib.engine import Engine
from lib.net_types import n_leaves
B,tau,eps,k=50,0.6,1e-6,4e-9
net=tiny_net('ac',k_cpt=k); eng=Engine(net,dry_run=True); plan=eng._plan(B,True,True)
nodes=eng.nodes; n=len(nodes)
rng=np.random.default_rng(13)
Rs={nd.idx:rng.standard_normal((B,len(nd.kids))).astype(np.float32) for nd in eng.switches}
ce={nd.idx:(rng.random(B)*3).astype(np.float32) for nd in eng.regs}
parent=plan.t_parent.numpy(); sink=plan.t_sink.numpy(); nsinks=plan.t_nsinks.numpy(); child=plan.t_child.numpy()
floor=plan.t_floor.numpy(); sw=plan.t_sw.numpy(); ops=plan.t_ops.numpy(); err=plan.t_err.numpy()
print('parent',parent,'sink',sink,'nsinks',nsinks,'sw',sw,'err',err,'child',child[:, :3].tolist())
Rtab=[Rs[nd.idx] for nd in eng.switches]; cetab=[ce[nd.idx] for nd in eng.regs]
def softmax(r):
    x=r/tau; e=np.exp(x-x.max()); return e/e.sum()
# forward p_tr
p_tr=np.ones((n,B))
for i in range(1,n):
    par=parent[i]
    for b in range(B):
        pt=p_tr[par,b]
        if nsinks[par]>=2:
            sm=softmax(Rtab[sw[par]][b]); pt=(pt-eps*floor[par])*sm[sink[i]]+eps*floor[i]
        p_tr[i,b]=pt
dR=[np.zeros_like(r) for r in Rtab]
for b in range(B):
    gp=np.zeros(n)
    for i in range(n):
        c=cetab[err[i]][b] if err[i]>=0 else 0.0
        gp[i]=(c+k*ops[i])/B
    for i in range(n-1,0,-1):
        par=parent[i]; g=gp[i]
        if nsinks[par]>=2: g*=softmax(Rtab[sw[par]][b])[sink[i]]
        gp[par]+=g
    for i in range(n):
        ns=nsinks[i]
        if ns<2: continue
        r=Rtab[sw[i]][b]; sm=softmax(r); pt=p_tr[i,b]
        gs=np.array([gp[child[i,j]]*(pt-eps*floor[i]) for j in range(ns)]); dot=(sm*gs).sum()
        dR[sw[i]][b]=sm*(gs-dot)/tau+pt*0.01*2*r/B
# autograd reference
Rt={i:torch.tensor(r,dtype=torch.float64,requires_grad=True) for i,r in Rs.items()}
fl=[eps*n_leaves(nd.layer)/n_leaves(net.root) for nd in nodes]
P=[None]*n; P[0]=torch.ones(B,dtype=torch.float64)
for nd in nodes[1:]:
    par=nodes[nd.parent]
    if len(par.kids)<2: P[nd.idx]=P[par.idx]
    else: P[nd.idx]=(P[par.idx]-fl[par.idx])*torch.softmax(Rt[par.idx]/tau,1)[:,nd.sink_idx]+fl[nd.idx]
tot=sum(P[nd.idx]*((torch.tensor(ce[nd.idx],dtype=torch.float64) if nd.idx in ce else 0)+k*float(ops[nd.idx])) for nd in nodes)
tot=tot+sum(P[nd.idx].detach()*0.01*(Rt[nd.idx]**2).sum(1) for nd in eng.switches)
tot.mean().backward()
for s,nd in enumerate(eng.switches): print(nd.idx, rel_err(dR[s],Rt[nd.idx].grad.numpy()), dR[s][0], Rt[nd.idx].grad.numpy()[0])
print('--- hypotheses for b=0,1')
for b in (0,1):
    gp=np.zeros(n)
    for i in range(n):
        c=cetab[err[i]][b] if err[i]>=0 else 0.0
        gp[i]=(c+k*ops[i])/B
    gp0=gp.copy()
    for i in range(n-1,0,-1):
        par=parent[i]; g=gp[i]
        if nsinks[par]>=2: g*=softmax(Rtab[sw[par]][b])[sink[i]]
        gp[par]+=g
    i=1; r=Rtab[0][b]; sm=softmax(r); pt=p_tr[i,b]
    gs=np.array([gp[child[i,j]]*pt for j in range(2)]); gs0=np.array([gp0[child[i,j]]*pt for j in range(2)])
    cdec=pt*0.01*2*r/B
    print('b',b,'r',r,'sm',sm,'gs',gs,'gs_noacc',gs0,'cdec',cdec)
    print('  true',sm*(gs-(sm*gs).sum())/tau+cdec,' dot0',sm*gs/tau+cdec, ' nosm', (gs-(sm*gs).sum())/tau+cdec, 'gs/tau+cdec', gs/tau+cdec)
