import torch, sys
sys.path.insert(0, '.')
from articulatory_b200._lib import call, ptr, F32, BF16
n = 70_710_000
a = torch.randn(n, device='cuda'); w = torch.empty(n, dtype=torch.bfloat16, device='cuda')
big = torch.empty(300_000_000, device='cuda')
def t(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
print("torch fp32->bf16 copy_", t(lambda: w.copy_(a)))
print("torch bf16->fp32 copy_", t(lambda: a.copy_(w)))
print("artic_cast fp32->bf16 ", t(lambda: call("artic_cast", ptr(a), F32, ptr(w), BF16, n)))
print("artic_cast bf16->fp32 ", t(lambda: call("artic_cast", ptr(w), BF16, ptr(a), F32, n)))
