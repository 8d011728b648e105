import torch
x = torch.empty(2*1024**3, dtype=torch.bfloat16, device='cuda')   # 4 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms = t(lambda: x.zero_()); print('fill   4 GiB: %.3f ms = %.2f TB/s (write only)' % (ms, x.numel()*2/ms/1e9))
ms = t(lambda: y.copy_(x)); print('copy   4 GiB: %.3f ms = %.2f TB/s (read+write)' % (ms, 2*x.numel()*2/ms/1e9))
ms = t(lambda: x.sum()); print('reduce 4 GiB: %.3f ms = %.2f TB/s (read only)' % (ms, x.numel()*2/ms/1e9))
