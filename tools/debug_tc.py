"""Debug helper: tcgen05 encoder attention forward vs the mma.sync kernel on a PACKED qkv activation at several batch sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, cabi
from vqacl_b200.engine import rel_bucket_table
H, S = 12, 56
tc = os.environ.get("VQACL_ATTN_TC") == "1"
for B in (8, 40, 110, 200, 320):
    torch.manual_seed(B)
    qkv = (torch.randn(B * S, 3 * 768, device="cuda") * 0.3).bfloat16()
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    table = torch.randn(32, H, device="cuda") * 0.5
    km = torch.zeros(B, S, device="cuda"); km[1, 9:20] = -10000.0
    o, lse = cabi.attention_fwd(q, k, v, B, H, S, S, rel_table=table, rel_bucket=rel_bucket_table(True), rel_mode=1, Lt=20, keymask=km)
    torch.cuda.synchronize()
    torch.save((o.cpu(), lse.cpu()), f"/tmp/tc_{int(tc)}_{B}.pt")
    other = f"/tmp/tc_{int(not tc)}_{B}.pt"
    if os.path.exists(other):
        o2, l2 = torch.load(other)
        d = (o.cpu().float() - o2.float()).abs()
        bad_rows = (d.max(dim=1).values > 0.05).nonzero().flatten()
        print(f"B={B}: max |dO| {d.max().item():.4f}  max |dlse| {(lse.cpu() - l2).abs().max().item():.4f}  bad rows {bad_rows.numel()}",
              (bad_rows[:8] // S).tolist(), (bad_rows[:8] % S).tolist())
