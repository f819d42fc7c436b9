"""Per-item timeline of CTA 0 of the tcgen05 attention backward, from clock64() stamps (needs a trace build:
VQACL_NVCC_EXTRA=-DVQ_ATTN_TRACE python -m vqacl_b200.build --force). Times in us relative to the producer's first stamp,
at the SM clock read from nvidia-smi.   VQACL_ATTN_TC_BWD=1 python tools/attn_trace.py"""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, cabi
from vqacl_b200._lib import lib, check
from vqacl_b200.engine import rel_bucket_table
B, H, S = 320, 12, 56
qkv = (torch.randn(B * S, 3 * 768, device="cuda") * 0.3).bfloat16()
q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
table = torch.randn(32, H, device="cuda") * 0.5
km = torch.zeros(B, S, device="cuda")
kw = dict(rel_table=table, rel_bucket=rel_bucket_table(True), rel_mode=1, Lt=20, keymask=km, causal=0)
dO = torch.randn(B * S, 768, device="cuda").bfloat16()
o, lse = cabi.attention_fwd(q, k, v, B, H, S, S, **kw)
for _ in range(3): cabi.attention_bwd(q, k, v, dO, lse, B, H, S, S, **kw)
torch.cuda.synchronize()
buf = torch.zeros(4 * 16 * 8, device="cuda", dtype=torch.int64)
L = lib()
L.vqacl_debug_attn_trace.argtypes = [ctypes.c_void_p]
check(L.vqacl_debug_attn_trace(buf.data_ptr()))
cabi.attention_bwd(q, k, v, dO, lse, B, H, S, S, **kw)
torch.cuda.synchronize()
check(L.vqacl_debug_attn_trace(None))
mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0])
mhz = mhz if mhz > 1500 else 1900.0
t = buf.cpu().view(4, 16, 8)
t0 = int(t[t > 0].min())
us = lambda x: (int(x) - t0) / mhz if int(x) > 0 else float("nan")
names = {0: ("producer", ["wait empty", "stage free", "TMA issued (all 8)", "header in smem", "first TMA", "after load 2", "after load 4", "after load 6"]),
         1: ("issuer", ["wait full", "full seen", "S/dP issued", "wait pfull", "pfull seen", "dV/dK/dQ issued"]),
         2: ("row warp 0", ["wait full", "full seen", "sfull seen", "D exchanged", "ofull(n-1) seen", "epilogue(n-1) done", "P/dS written", "pfull arrive"]),
         3: ("row warp 15", ["wait full", "full seen", "sfull seen", "D exchanged", "ofull(n-1) seen", "epilogue(n-1) done", "P/dS written", "pfull arrive"])}
print(f"SM clock {mhz:.0f} MHz (idle reading; the stamps are cycles / this)")
for role, (nm, slots) in names.items():
    print(f"--- {nm}: " + " | ".join(slots))
    for n in range(14):
        if int(t[role, n].max()) == 0: continue
        print(f"  item {n:2d}: " + " ".join(f"{us(t[role, n, i]):7.2f}" for i in range(len(slots))))
