import torch
dev=torch.device("cuda:0")
x=torch.empty(1<<30, dtype=torch.uint8, device=dev)   # 1 GiB
y=torch.empty(1<<30, dtype=torch.uint8, device=dev)
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
ms=t(lambda: x.zero_()); print("write-only  %.0f GB/s"%((1<<30)/ms/1e6))
ms=t(lambda: x.view(torch.float32).sum()); print("read-only   %.0f GB/s"%((1<<30)/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy (r+w)  %.0f GB/s"%(2*(1<<30)/ms/1e6))
