"""GPU bring-up check of the tcgen05 GEMM through the C-ABI (vqacl_gemm_bf16) against torch fp32 matmul
of the same bf16-rounded operands. Run on the GPU box: python tools/gemm_check.py [--bench]."""
import ctypes
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqacl_b200._lib import lib, check, ptr, cur_stream

EPI_BF16, EPI_RELU, EPI_RESID, EPI_ATOMIC, EPI_RELUBWD, EPI_F32 = 0, 1, 2, 3, 4, 5


def gemm(A, a_mn, B, b_mn, C, R, M, N, K, epi, alpha=1.0, splits=1, bn=0):
    L = lib()
    check(L.vqacl_gemm_bf16(ptr(A), A.stride(0), int(a_mn), ptr(B), B.stride(0), int(b_mn), ptr(C), C.stride(0),
                            ptr(R), R.stride(0) if R is not None else 0, M, N, K, epi, ctypes.c_float(alpha),
                            splits, bn, cur_stream()))


def pack_bits(bits):
    M, N = bits.shape
    W = (N + 31) // 32
    b = torch.zeros(M, W * 32, device=bits.device, dtype=torch.int64)
    b[:, :N] = bits
    v = (b.view(M, W, 32) << torch.arange(32, device=bits.device)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32)


def run_case(M, N, K, a_mn, b_mn, epi, bn, splits=1):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + bn)
    Al = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    Bl = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    A = Al.t().contiguous() if a_mn else Al
    B = Bl.t().contiguous() if b_mn else Bl
    ref = Al.float() @ Bl.float().t()
    R = None
    if epi in (EPI_BF16, EPI_RELU, EPI_RELUBWD):
        C = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    else:
        C = torch.full((M, N), 0.5, device="cuda", dtype=torch.float32)
    if epi == EPI_RELU:
        ref = ref.relu()
    elif epi == EPI_RESID:
        R = torch.randn(M, N, device="cuda", generator=g)
        ref = ref + R
    elif epi == EPI_ATOMIC:
        ref = ref + 0.5
    elif epi == EPI_RELUBWD:
        keep = torch.rand(M, N, device="cuda", generator=g) > 0.5             # ReLU sign bitmask, bit i of word w = column 32w+i
        R = pack_bits(keep)
        ref = ref * keep
    gemm(A, a_mn, B, b_mn, C, R, M, N, K, epi, 1.0, splits, bn)
    torch.cuda.synchronize()
    err = (C.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    tol = 1e-2 if C.dtype == torch.bfloat16 else 2e-3
    ok = err / scale < tol
    print(f"{'OK ' if ok else 'BAD'} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} bn={bn} splits={splits} "
          f"max_err={err:.4g} rel={err / scale:.3g}", flush=True)
    return ok


def bench(M, N, K, a_mn, b_mn, epi, bn, splits=1, iters=20):
    Al = torch.randn(M, K, device="cuda").bfloat16()
    Bl = torch.randn(N, K, device="cuda").bfloat16()
    A = Al.t().contiguous() if a_mn else Al
    B = Bl.t().contiguous() if b_mn else Bl
    C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi in (0, 1, 4) else torch.float32)
    R = torch.zeros(M, N, device="cuda", dtype=torch.float32) if epi == EPI_RESID else (torch.full((M, (N + 31) // 32), -1, device="cuda", dtype=torch.int32) if epi == EPI_RELUBWD else None)
    for _ in range(3):
        gemm(A, a_mn, B, b_mn, C, R, M, N, K, epi, 1.0, splits, bn)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gemm(A, a_mn, B, b_mn, C, R, M, N, K, epi, 1.0, splits, bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS reference for context
    Cc = Al @ Bl.t()
    e0.record()
    for _ in range(iters):
        Cc = Al @ Bl.t()
    e1.record()
    torch.cuda.synchronize()
    msc = e0.elapsed_time(e1) / iters
    print(f"bench M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} bn={bn} splits={splits}: {ms * 1e3:.1f} us "
          f"{tf:.1f} TFLOP/s | cuBLAS {msc * 1e3:.1f} us {2.0 * M * N * K / msc / 1e9:.1f} TFLOP/s", flush=True)


def main():
    ok = True
    # smallest first: a wrong descriptor shows up here
    for bn in (64, 128, 256):
        ok &= run_case(128, bn, 64, False, False, EPI_F32, bn)
        ok &= run_case(128, bn, 256, False, False, EPI_F32, bn)
        ok &= run_case(128, bn, 256, False, True, EPI_F32, bn)
        ok &= run_case(128, bn, 256, True, True, EPI_F32, bn)
    for (a_mn, b_mn) in ((False, False), (False, True), (True, True)):
        for bn in (64, 128, 256):
            ok &= run_case(1000, 776, 520, a_mn, b_mn, EPI_F32, bn)
            ok &= run_case(1600, 768, 768, a_mn, b_mn, EPI_BF16, bn)
    for epi in (EPI_RELU, EPI_RESID, EPI_ATOMIC, EPI_RELUBWD):
        ok &= run_case(17920 // 4, 768, 768, False, False, epi, 0)
    ok &= run_case(768, 768, 17920, True, True, EPI_ATOMIC, 0, splits=8)
    ok &= run_case(3072, 768, 4480, True, True, EPI_ATOMIC, 256, splits=5)
    ok &= run_case(1600, 32200, 768, False, False, EPI_BF16, 0)
    ok &= run_case(1600, 768, 32200, False, True, EPI_BF16, 0)
    ok &= run_case(32200, 768, 1600, True, True, EPI_ATOMIC, 0)
    # CTA-pair kernel (bn = 512: 256 x 256 tiles on clusters of two CTAs)
    if "--no-pair" not in sys.argv:
        ok &= run_case(256, 256, 64, False, False, EPI_F32, 512)
        ok &= run_case(256, 256, 256, False, False, EPI_F32, 512)
        ok &= run_case(256, 256, 256, False, True, EPI_F32, 512)
        ok &= run_case(256, 256, 256, True, True, EPI_F32, 512)
        for (a_mn, b_mn) in ((False, False), (False, True), (True, True)):
            ok &= run_case(1000, 776, 520, a_mn, b_mn, EPI_F32, 512)
            ok &= run_case(4480, 2304, 768, a_mn, b_mn, EPI_BF16, 512)
        for epi in (EPI_RELU, EPI_RESID, EPI_ATOMIC, EPI_RELUBWD):
            ok &= run_case(17920 // 4, 768, 768, False, False, epi, 512)
        ok &= run_case(3072, 768, 4480, True, True, EPI_ATOMIC, 512, splits=5)
    print("ALL OK" if ok else "FAILURES", flush=True)
    if "--bench" in sys.argv:
        bench(17920, 2304, 768, False, False, EPI_BF16, 0)
        bench(17920, 768, 768, False, False, EPI_RESID, 0)
        bench(17920, 3072, 768, False, False, EPI_RELU, 0)
        bench(17920, 768, 3072, False, False, EPI_RESID, 0)
        bench(17920, 768, 3072, False, True, EPI_BF16, 0)
        bench(3072, 768, 17920, True, True, EPI_ATOMIC, 0, splits=4)
        bench(768, 768, 17920, True, True, EPI_ATOMIC, 0, splits=8)
        bench(8192, 8192, 8192, False, False, EPI_BF16, 256)
        if "--no-pair" not in sys.argv:
            bench(17920, 2304, 768, False, False, EPI_BF16, 512)
            bench(17920, 768, 768, False, False, EPI_RESID, 512)
            bench(17920, 3072, 768, False, False, EPI_RELU, 512)
            bench(17920, 768, 3072, False, False, EPI_RESID, 512)
            bench(17920, 768, 3072, False, True, EPI_BF16, 512)
            bench(17920, 3072, 768, False, True, EPI_RELUBWD, 512)
            bench(3072, 768, 17920, True, True, EPI_ATOMIC, 512, splits=4)
            bench(8192, 8192, 8192, False, False, EPI_BF16, 512)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
