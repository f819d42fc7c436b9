"""configs[4]: eval-only greedy answer generation, batch 512 (KV-cached native loop): samples/s on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch, vqacl_b200 as V, vlt5_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=0.0)).to("cuda"); m.eval()
g = torch.Generator().manual_seed(1)
m.Q_prototype = torch.randn(10, 768, generator=g); m.V_prototype = torch.randn(80, 768, generator=g)
b = {k: v.cuda() for k, v in O.synthetic_batch(B, seed=3).items()}
for _ in range(2): out = m.test_step(b)["token_ids"]
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 5
for _ in range(K): out = m.test_step(b)["token_ids"]
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print(f"greedy decode B={B}: {out.shape[1]} tokens, {1e3*dt:.1f} ms per batch, {B/dt:.0f} samples/s")
