import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, cabi
n = 224_540_160
p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda"); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
pb = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
for _ in range(3): cabi.adamw(p, g, m, v, n, 1e-4, 0.9, 0.999, 1e-6, 0.01, 1, None, 0.0, pb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): cabi.adamw(p, g, m, v, n, 1e-4, 0.9, 0.999, 1e-6, 0.01, 2, None, 0.0, pb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"adamw {n/1e6:.1f} M params: {ms*1e3:.0f} us, {n*30/ms/1e6:.0f} GB/s")
