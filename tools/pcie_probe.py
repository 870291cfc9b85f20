import torch,time
n=83*1000*1000
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(33*1000*1000,dtype=torch.uint8).pin_memory(); d2=torch.empty(33*1000*1000,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,k=10):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/k
a=t(lambda: d.copy_(h,non_blocking=True)); print('H2D 83MB %.3f ms %.1f GB/s'%(a*1e3,n/a/1e9))
b=t(lambda: h2.copy_(d2,non_blocking=True)); print('D2H 33MB %.3f ms %.1f GB/s'%(b*1e3,33e6/b/1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
c=t(both); print('both %.3f ms'%(c*1e3))
