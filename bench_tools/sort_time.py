import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arvae_b200
for B in (8192, 65536, 262144):
    a = torch.randn(B, device="cuda")
    for _ in range(3): arvae_b200.attr_argsort(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): arvae_b200.attr_argsort(a)
    e1.record(); torch.cuda.synchronize()
    print(B, "argsort (1 dim) us:", round(e0.elapsed_time(e1) / 50 * 1e3, 1))
