import torch, time, numpy as np, sys
sys.path.insert(0,'.')
x=torch.empty(1300*1024*1024//8, dtype=torch.float64, device='cuda')
h=torch.empty_like(x, device='cpu').pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('D2H GB/s', x.numel()*8/dt/1e9)
    torch.cuda.synchronize(); t=time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D GB/s', x.numel()*8/dt/1e9)
del x,h
from romhighcontrast_b200.engine import Engine
import bench
eng=Engine((4,4),64)
K=10000
y=bench.sample_params(K,42)
U=torch.empty((K,eng.D),dtype=torch.float64).pin_memory().numpy()
yp=np.ascontiguousarray(y.reshape(K,-1))
eng.generate_solutions_host(yp,out=U)
for ws in [4, 6, 8, 12]:
    eng.set_option("host_chunks", ws)
    eng.generate_solutions_host(yp,out=U)
    torch.cuda.synchronize(); t=time.perf_counter(); eng.generate_solutions_host(yp,out=U); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('host_chunks',ws,'e2e ms',dt*1e3, K/dt)
Upg=np.empty((K,eng.D))
eng.set_option("host_chunks", 4)
for rep in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); eng.generate_solutions_host(yp,out=Upg); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('pageable destination e2e ms',dt*1e3, K/dt, 'equal to pinned result:', bool(np.array_equal(Upg,U)))
yd=eng.params(y); x=eng.empty(K,eng.Dp)
eng.set_option("workspace_gb", 48)
for k in [10000,2500,1250]:
    eng.solve(yd[:k],out=x[:k]); torch.cuda.synchronize(); t=time.perf_counter(); eng.solve(yd[:k],out=x[:k]); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('resident K',k,'ms',dt*1e3,k/dt)
