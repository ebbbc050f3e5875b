import torch,time
x=torch.empty(128*4194304,dtype=torch.float64,device='cuda')
y=torch.empty(128*4194304//2,dtype=torch.float64,device='cuda')
for name,fn in [("fill",lambda: x.fill_(1.0)),("copy",lambda: x[:y.numel()].copy_(y)),("scale2x", lambda: torch.mul(y,2.0,out=x[:y.numel()]))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    nbytes = x.numel()*8 if name=="fill" else y.numel()*16
    print(name, ms, "ms", nbytes/ms/1e6, "GB/s")
