"""Do the (tensor-bound) GEMM and the (HBM-bound) AdamW kernel overlap when issued on two streams?"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch, cabi
from vqacl_b200._lib import lib, check, ptr
L = cabi._L()
M = N = K = 8192
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16(); C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
n = 224_000_000
p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda"); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
pb = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def gemm(st, bn):
    check(L.vqacl_gemm_bf16(ptr(A), K, 0, ptr(B), K, 0, ptr(C), N, None, 0, M, N, K, 0, ctypes.c_float(1.0), 1, bn, ctypes.c_void_p(st.cuda_stream)))
def adam(st):
    check(L.vqacl_adamw_hf(ptr(p), ptr(g), ptr(m), ptr(v), ptr(pb), n, n, 1e-4, 0.9, 0.999, 1e-6, 0.01, 1, None, 0.0, ctypes.c_void_p(st.cuda_stream)))
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
for bn in (256, 512):
    tg = timeit(lambda: gemm(s1, bn)); ta = timeit(lambda: adam(s2))
    def both():
        gemm(s1, bn); gemm(s1, bn); adam(s2)
    tb = timeit(both)
    print(f"bn={bn}: gemm {tg:.3f} ms, adamw {ta:.3f} ms, 2 gemm || adamw {tb:.3f} ms (serial would be {2*tg+ta:.3f})")
