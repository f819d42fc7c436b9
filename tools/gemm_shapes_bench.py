"""Every GEMM shape of the configs[1] step, auto tile choice vs forced choices vs cuBLAS (bf16), interleaved so that clocks
affect all candidates alike. python tools/gemm_shapes_bench.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqacl_b200._lib import lib, check, ptr, cur_stream
L = lib()
B = 320
M, Md, M2 = B * 56, B * 5, B * 58
# name, M, N, K, a_mn, b_mn, epi, splits, count per step
S = [("enc qkv", M, 2304, 768, 0, 0, 0, 1, 12), ("enc o (+res)", M, 768, 768, 0, 0, 2, 1, 12), ("enc wi (+relu)", M, 3072, 768, 0, 0, 1, 1, 12),
     ("enc wo (+res)", M, 768, 3072, 0, 0, 2, 1, 12), ("enc dX wo (relubwd)", M, 3072, 768, 0, 1, 4, 1, 12), ("enc dX wi", M, 768, 3072, 0, 1, 0, 1, 12),
     ("enc dX o", M, 768, 768, 0, 1, 0, 1, 12), ("enc dX qkv", M, 768, 2304, 0, 1, 0, 1, 12),
     ("enc dW wo", 768, 3072, M, 1, 1, 3, 0, 12), ("enc dW wi", 3072, 768, M, 1, 1, 3, 0, 12), ("enc dW o", 768, 768, M, 1, 1, 3, 0, 12),
     ("enc dW qkv", 2304, 768, M, 1, 1, 3, 0, 12),
     ("cross KV fwd", M2, 18432, 768, 0, 0, 0, 1, 1), ("cross KV dX", M2, 768, 18432, 0, 1, 0, 1, 1), ("cross KV dW", 18432, 768, M2, 1, 1, 3, 0, 1),
     ("vis feat", B * 36, 768, 2048, 0, 0, 5, 1, 1), ("lm head", Md, 32200, 768, 0, 0, 0, 1, 1), ("lm dW", 32200, 768, Md, 1, 1, 3, 0, 1),
     ("dec qkv", Md, 2304, 768, 0, 0, 0, 1, 12), ("dec o", Md, 768, 768, 0, 0, 2, 1, 36), ("dec wi", Md, 3072, 768, 0, 0, 1, 1, 12),
     ("dec wo", Md, 768, 3072, 0, 0, 2, 1, 12), ("dec dX 768", Md, 768, 768, 0, 1, 0, 1, 36), ("dec dW 768", 768, 768, Md, 1, 1, 3, 0, 36),
     ("dec dW wi", 3072, 768, Md, 1, 1, 3, 0, 12)]
def splits_dw(n_out, n_in, rows):
    tiles = ((n_out + 127) // 128) * ((n_in + 255) // 256)
    kb = (rows + 63) // 64
    return max(1, min((2 * 148) // max(tiles, 1), kb // 8))
tot = {}
for name, m, n, k, amn, bmn, epi, sp, cnt in S:
    A = torch.randn((k, m) if amn else (m, k), device="cuda").bfloat16()
    Bm = torch.randn((k, n) if bmn else (n, k), device="cuda").bfloat16()
    f32 = epi in (2, 3, 5)
    ldc = (n + 255) // 256 * 256 if name == "lm head" else n
    C = torch.zeros(m, ldc, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    R = torch.zeros(m, n, device="cuda") if epi == 2 else (torch.full((m, (n + 31) // 32), -1, device="cuda", dtype=torch.int32) if epi in (1, 4) else None)
    if sp == 0: sp = splits_dw(m, n, k)
    def run(bn):
        def call():
            check(L.vqacl_gemm_bf16(ptr(A), A.stride(0), amn, ptr(Bm), Bm.stride(0), bmn, ptr(C), C.stride(0), ptr(R),
                                    R.stride(0) if R is not None else 0, m, n, k, epi, ctypes.c_float(1.0), sp, bn, cur_stream()))
        for _ in range(2): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): call()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 100
    Al = A.t() if amn else A
    Bl = Bm.t() if bmn else Bm
    def cublas():
        for _ in range(2): Al @ Bl.t()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): Al @ Bl.t()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 100
    res = {}
    for rep in range(2):
        for bn in (0, 64, 128, 256, 512):
            if bn == 512 and (m < 192 or n < 192): continue
            t = run(bn); res[bn] = min(res.get(bn, 1e9), t)
        t = cublas(); res["cublas"] = min(res.get("cublas", 1e9), t)
    best = min((v, k2) for k2, v in res.items() if k2 not in (0, "cublas"))
    print(f"{name:22s} M={m:6d} N={n:6d} K={k:6d} sp={sp:2d} | auto {res[0]:7.1f} | " + " ".join(f"{k2}:{v:7.1f}" for k2, v in res.items() if k2 != 0) +
          f" | best {best[1]} ({2.0*m*n*k/res[0]/1e6:6.0f} TF auto)")
    for k2, v in res.items(): tot[k2] = tot.get(k2, 0) + v * cnt if k2 in (0, "cublas") else 0
    tot["best"] = tot.get("best", 0) + best[0] * cnt
print("per-step totals (us): auto", round(tot[0]), "best-forced", round(tot["best"]), "cuBLAS (no epilogues)", round(tot["cublas"]))
