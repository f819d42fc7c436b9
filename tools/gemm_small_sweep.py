"""Where do the ~10 us of a decoder-sized GEMM (M = 1600) go? K sweep (fixed overhead vs per-k-block cost), tile widths, split-K.
python tools/gemm_small_sweep.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqacl_b200._lib import lib, check, ptr, cur_stream
L = lib()
def t(M, N, K, epi, bmn=0, bn=0, splits=1, reps=20):
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn((K, N) if bmn else (N, K), device="cuda").bfloat16()
    f32 = epi in (2, 3, 5)
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    R = torch.randn(M, N, device="cuda") if epi == 2 else None
    def call():
        check(L.vqacl_gemm_bf16(ptr(A), A.stride(0), 0, ptr(B), B.stride(0), bmn, ptr(C), C.stride(0), ptr(R), R.stride(0) if R is not None else 0,
                                M, N, K, epi, ctypes.c_float(1.0), splits, bn, cur_stream()))
    return graph_time(call, reps)
def graph_time(call, reps=20):
    """`reps` launches captured into one CUDA graph: the replay is not bound by the ~9 us a ctypes launch costs from Python"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): call()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): call()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        g.replay()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
def empty_launch():
    x = torch.zeros(32, device="cuda")
    return graph_time(lambda: x.add_(1))
print(f"tiny torch kernel, 20 in a graph: {empty_launch():.1f} us")
for M in (1600, 128):
    for N in (768, 2304):
        for bn in (64, 128, 256):
            print(f"M={M} N={N} bn={bn} epi=0 | " + " ".join(f"K={K}: {t(M, N, K, 0, bn=bn):5.1f}" for K in (64, 128, 256, 512, 768, 1536, 3072)))
for epi in (0, 2, 3):
    print(f"M=1600 N=768 K=768 epi={epi} | " + " ".join(f"bn={bn} sp={sp}: {t(1600, 768, 768, epi, bn=bn, splits=sp):5.1f}" for bn in (64, 128) for sp in ((1, 2, 3) if epi == 3 else (1,))))
