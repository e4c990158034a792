import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'multipath-nn_b200'), os.path.join(ROOT,'tests')]
import numpy as np, torch
from util import tiny_net
B,tau,eps,k=50,0.6,1e-6,4e-9
net=tiny_net('ac',k_cpt=k).configure(precision='fp32'); eng=net._get_engine(); plan=eng._plan(B,True,True)
rng=np.random.default_rng(13)
for nd in eng.switches:
    plan.rtr[nd.idx].R.copy_(torch.from_numpy(rng.standard_normal((B,len(nd.kids))).astype(np.float32)))
for nd in eng.regs:
    plan.reg[nd.idx].c_err.copy_(torch.from_numpy((rng.random(B)*3).astype(np.float32)))
hyp=np.zeros(8,np.float32); hyp[2]=tau; hyp[3]=eps; hyp[4]=k
eng.hyp.copy_(torch.from_numpy(hyp)); eng.stream=None
plan.fwd_ops[-1](); plan.bwd_ops[0](); torch.cuda.synchronize()
n=len(eng.nodes)
gp=plan.route_scratch[:n*B].reshape(n,B).cpu().numpy()
print('gp b=0', gp[:,0]); print('gp b=1', gp[:,1])
print('dR sw0 b0,b1', plan.rtr[1].dR[:2].cpu().numpy()); print('dR sw1 b0,b1', plan.rtr[3].dR[:2].cpu().numpy())
print('R sw0', plan.rtr[1].R[:2].cpu().numpy(), 'ptr', plan.p_tr[:,0].cpu().numpy())
print('tabs R', plan.R_tab.cpu().numpy(), 'dR', plan.dR_tab.cpu().numpy(), [plan.rtr[i].R.data_ptr() for i in (1,3)], [plan.rtr[i].dR.data_ptr() for i in (1,3)])
