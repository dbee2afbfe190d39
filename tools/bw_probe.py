import torch, time
x = torch.randn(1024, 1024, 1024, device="cuda")
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.sum())
print(f"sum: {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s read")
ms = t(lambda: x.max())
print(f"max: {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s read")
y = torch.empty_like(x)
ms = t(lambda: y.copy_(x))
print(f"copy: {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s r+w")
