"""RMSNorm backward alone at the encoder's and the decoder's row counts, with and without the norm-weight gradient (whose 768
global atomics per CTA all land on the same 768 addresses at the kernel's tail), replayed from a CUDA graph.
python tools/rms_bwd_probe.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqacl_b200._lib import lib, check, ptr, cur_stream
L = lib()
def graph_us(fn, reps=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for M in (17920, 1600):
    x = torch.randn(M, 768, device="cuda"); dn = torch.randn(M, 768, device="cuda").bfloat16(); w = torch.ones(768, device="cuda")
    g_in = torch.randn(M, 768, device="cuda"); g_out = torch.empty_like(g_in); gb = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(768, device="cuda")
    for use_dw in (True, False):
        def call():
            check(L.vqacl_rmsnorm_bwd(ptr(dn), ptr(x), ptr(w), ptr(g_in), ptr(g_out), ptr(gb), ptr(dw) if use_dw else None, M,
                                      ctypes.c_float(1e-6), ctypes.c_float(1.0), cur_stream()))
        print(f"rmsnorm_bwd M={M} dw={'yes' if use_dw else 'no '}: {graph_us(call):.1f} us")
