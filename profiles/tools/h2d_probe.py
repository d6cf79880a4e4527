import torch, time
x = torch.empty(83_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty(83_000_000, dtype=torch.uint8, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/10
print("pinned H2D 83 MB: %.2f ms = %.1f GB/s" % (dt*1e3, 83e6/dt/1e9))
import numpy as np
a = np.random.rand(83_000_000//8); b = np.empty_like(a)
t0=time.perf_counter(); 
for _ in range(5): np.copyto(b,a)
dt=(time.perf_counter()-t0)/5
print("single-thread memcpy 83 MB: %.2f ms = %.1f GB/s" % (dt*1e3, 83e6/dt/1e9))
