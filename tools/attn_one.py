"""Launch the encoder-shaped attention fwd/bwd a few times (for ncu / timing). python tools/attn_one.py [B] [mode]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, cabi
from vqacl_b200.engine import rel_bucket_table
B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
mode = sys.argv[2] if len(sys.argv) > 2 else "enc"
H = 12
Sq, Sk = {"enc": (56, 56), "dec": (5, 5), "cross": (5, 58)}[mode]
if mode == "enc":
    qkv = (torch.randn(B * Sq, 3 * 768, device="cuda") * 0.3).bfloat16()
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
elif mode == "dec":
    qkv = (torch.randn(B * Sq, 3 * 768, device="cuda") * 0.3).bfloat16()
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
else:
    q = (torch.randn(B * Sq, 768, device="cuda") * 0.3).bfloat16()
    kvw = int(os.environ.get("KVW", 24 * 768))
    kv = (torch.randn(B * Sk, kvw, device="cuda") * 0.3).bfloat16()
    k, v = kv[:, :768], kv[:, 768:1536]
table = torch.randn(32, H, device="cuda") * 0.5
bucket = rel_bucket_table(mode == "enc")
km = torch.zeros(B, Sk, device="cuda")
kw = dict(rel_table=table, rel_bucket=bucket, rel_mode=1 if mode == "enc" else 2, Lt=20, keymask=km, causal=int(mode == "dec")) if mode != "cross" else dict(keymask=km)
dO = torch.randn(B * Sq, 768, device="cuda").bfloat16()
def fwd(): return cabi.attention_fwd(q, k, v, B, H, Sq, Sk, **kw)
o, lse = fwd()
def bwd(): return cabi.attention_bwd(q, k, v, dO, lse, B, H, Sq, Sk, o_saved=(o if mode == 'cross' else None), **kw)
def graph_us(fn, reps=20):
    """`reps` launches replayed from one CUDA graph: a ctypes launch costs ~9 us from Python, more than the small kernels take"""
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
for fn, nm in ((fwd, "fwd"), (bwd, "bwd")):
    if os.environ.get("GRAPH", "0") == "1":
        print(f"attn {mode} {nm} B={B}: {graph_us(fn):.1f} us (graph replay)")
        continue
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"attn {mode} {nm} B={B}: {e0.elapsed_time(e1) * 100:.1f} us")
