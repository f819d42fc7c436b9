"""Where does the end-to-end step lose time vs the device-resident step? (a) resident, no sync; (b) resident + .item();
(c) prefetcher, no sync; (d) prefetcher + .item(); (e) plain pinned batch through train_step's own H2D + .item()."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch, vqacl_b200 as V, vlt5_oracle as O
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=0.1)).to("cuda"); m.train()
opt = V.FusedAdamW(m, overlap_with_next_forward=True)
host = [{k: v.pin_memory() for k, v in O.synthetic_batch(320, seed=i, task_id=3).items()} for i in range(4)]
dev = [{k: v.cuda() for k, v in b.items()} for b in host]
def step(b, sync):
    r = m.train_step(b, 3, 0.5, 0.3); r["loss"].backward(); opt.step(max_grad_norm=5.0); opt.zero_grad()
    if sync: return r["loss"].item()
def run(name, gen, sync, K=20):
    for b in gen(4): step(b, sync)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for b in gen(K): step(b, sync)
    m.param_sync(); e1.record(); torch.cuda.synchronize()
    print(f"{name:45s} {e0.elapsed_time(e1)/K:7.3f} ms/step (wall {1e3*(time.perf_counter()-t0)/K:7.3f})")
res = lambda K: (dev[i % 4] for i in range(K))
pre = lambda K: V.BatchPrefetcher((host[i % 4] for i in range(K)), "cuda")
pin = lambda K: (host[i % 4] for i in range(K))
run("(a) resident, no sync", res, False)
run("(b) resident + loss.item()", res, True)
run("(c) prefetcher, no sync", pre, False)
run("(d) prefetcher + loss.item()", pre, True)
run("(e) pinned batch via train_step H2D + item", pin, True)
run("(a) resident, no sync", res, False)
