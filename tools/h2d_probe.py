import torch, time
m,n=1000000,500
host=torch.empty((m,n),dtype=torch.int32,pin_memory=True); host.random_(1,5)
dev=torch.empty((m,n),dtype=torch.int32,device='cuda')
for k in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
    print("full 2GB H2D ms", (time.perf_counter()-t)*1e3)
for k in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    dev[:m//2].copy_(host[:m//2], non_blocking=True); torch.cuda.synchronize()
    print("half 1GB H2D ms", (time.perf_counter()-t)*1e3)
print(bool((dev[:m//2].cpu()==host[:m//2]).all()))
