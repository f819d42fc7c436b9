"""Launch one GEMM shape a few times (for ncu). python tools/gemm_one.py M N K epi [a_mn b_mn bn splits reps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqacl_b200._lib import lib, check, ptr, cur_stream
a = [int(x) for x in sys.argv[1:]]
M, N, K, epi = a[:4]
amn, bmn, bn, splits, reps = (a[4:] + [0, 0, 0, 1, 5][len(a) - 4:])
A = torch.randn((K, M) if amn else (M, K), device="cuda").bfloat16()
B = torch.randn((K, N) if bmn else (N, K), device="cuda").bfloat16()
f32 = epi in (2, 3, 5)
C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
R = torch.randn(M, N, device="cuda") if epi == 2 else (torch.randint(-2**31, 2**31 - 1, (M, (N + 31) // 32), device="cuda", dtype=torch.int32) if epi == 4 else None)
L = lib()
def call():
    check(L.vqacl_gemm_bf16(ptr(A), A.stride(0), amn, ptr(B), B.stride(0), bmn, ptr(C), C.stride(0), ptr(R),
                            R.stride(0) if R is not None else 0, M, N, K, epi, ctypes.c_float(1.0), splits, bn, cur_stream()))
for _ in range(3): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): call()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print(f"M={M} N={N} K={K} epi={epi} amn={amn} bmn={bmn} bn={bn} splits={splits}: {us:.1f} us {2*M*N*K/us/1e6:.1f} TFLOP/s")
