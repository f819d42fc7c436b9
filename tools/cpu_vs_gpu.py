"""Is the step CPU-launch-bound? wall time to ISSUE K steps vs time until the GPU finishes them."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch, vqacl_b200 as V, vlt5_oracle as O
ov = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=0.1)).to("cuda"); m.train()
opt = V.FusedAdamW(m, overlap_with_next_forward=bool(ov))
b = {k: v.cuda() for k, v in O.synthetic_batch(320, task_id=3).items()}
def step():
    r = m.train_step(b, 3, 0.5, 0.3); r["loss"].backward(); opt.step(max_grad_norm=5.0); opt.zero_grad()
for _ in range(5): step()
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
for _ in range(K): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"overlap={ov}: issue {1e3*(t1-t0)/K:.2f} ms/step, complete {1e3*(t2-t0)/K:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
