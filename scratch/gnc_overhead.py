import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import psulvsb_b200
from psulvsb_b200 import capi, stages, synth
from oracle import oracle as O
pair = synth.make_pair(5000, 0.95, 3)
pi, pj = O.reduced_set(pair['src'], pair['dst'], 0.1)
rng = np.random.default_rng(0)
d_src, d_dst = stages.to_device_points(pair['src']), stages.to_device_points(pair['dst'])
for K in [64, 1024, 4096, 20000, 60000]:
    sel = rng.permutation(len(pi))[:K]
    e = torch.from_numpy(np.stack([pi[sel], pj[sel]], axis=1).astype(np.int32)).cuda()
    for _ in range(3):
        R, inl, its, cost, n_inl = stages.gnc_tls_rotation(d_src, d_dst, e, 0.1, 100, 1.4, 0.005)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        w = torch.zeros(K, dtype=torch.float64, device='cuda')
        Rr = torch.zeros(9, dtype=torch.float64, device='cuda'); info = torch.zeros(4, dtype=torch.int32, device='cuda'); c = torch.zeros(1, dtype=torch.float64, device='cuda')
        capi.check(capi.lib().psulvsb_gnc_tls_rotation(torch.cuda.current_stream().cuda_stream, d_src.data_ptr(), d_dst.data_ptr(), e.data_ptr(), K, 1.0, 0.1, 100, 1.4, 0.005, None, w.data_ptr(), Rr.data_ptr(), None, info.data_ptr(), c.data_ptr()))
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)/10
    print(f"K={K:6d} its={its:3d} ms={ms:.4f} us/iter={1000*ms/its:.2f}")
