"""Read-only / write-only / copy HBM bandwidth with torch built-ins (context for the write-heavy kernels)."""
import torch, time
dev=torch.device('cuda',0)
n=1<<30
x=torch.empty(n,dtype=torch.float32,device=dev)
y=torch.empty(n,dtype=torch.float32,device=dev)
def t(f,reps=10):
    f(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
ms=t(lambda: x.zero_()); print('zero_ (write only) GB/s', 4*n/ms/1e6)
ms=t(lambda: x.fill_(1.5)); print('fill_ GB/s', 4*n/ms/1e6)
ms=t(lambda: y.copy_(x)); print('copy (r+w) GB/s', 8*n/ms/1e6)
ms=t(lambda: x.sum()); print('sum (read only) GB/s', 4*n/ms/1e6)
