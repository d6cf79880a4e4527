import sys, numpy as np
sys.path.insert(0,'/root/repo')
import psulvsb_b200
from psulvsb_b200 import capi
from oracle import oracle as O
b=np.load('/root/repo/tests/golden/benchmark.npz')
k=1
src,dst=b[f'b{k}_src'],b[f'b{k}_dst']; nb=float(b[f'b{k}_noise_bound'][0])
kw=dict(noise_bound=nb,cbar2=1.0,estimate_scaling=1,rotation_cost_threshold=0.005,wallclock_cap_s=0.0,inloop_noise_bound=nb,score_noise_bound=nb)
so,to=O.solve(O.default_params(seed=1,**kw),src,dst)
h=capi.Handle(0)
sg,tg=h.solve(capi.default_params(seed=1,**kw),capi.HostProblem(src,dst),trace_cap=4096)
F=["host_round","local_iter","n_sampled_lines","n_sampled_points","basic_choose","gnc_iterations","rot_inliers","n_rot_points","similar","curr_count","best_count","local_r"]
for i in range(max(len(to['local']),len(tg['local']))):
    for name,tr in (('O',to),('G',tg)):
        if i<len(tr['local']):
            a=tr['local'][i]
            print(name,[getattr(a,f) for f in F], round(a.p_local,4), a.scale, np.round(np.array(a.t[:]),5), np.round(np.array(a.R[:]),4)[:4])
    if i>6: break
for name,tr in (('O',to),('G',tg)):
    for hh in tr['host']: print(name,'H',hh.host_round,hh.curr_count,hh.best_host,hh.host_r,hh.p_host)
