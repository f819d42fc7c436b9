"""Cost of running backward stage by stage (the multi-GPU path) vs in one call, on ONE GPU with a 1-rank NCCL group:
separates 'staging + event + side-stream join overhead' from actual communication cost."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch, torch.distributed as dist, vqacl_b200 as V, vlt5_oracle as O
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=0.1)).to("cuda"); m.train()
opt = V.FusedAdamW(m, overlap_with_next_forward=True)
b = {k: v.cuda() for k, v in O.synthetic_batch(320, task_id=3).items()}
def run(staged, allreduce, K=20):
    m._world = (lambda: 2) if staged else (lambda: 1)
    if staged and not allreduce:
        dist_all_reduce = dist.all_reduce
        dist.all_reduce = lambda *a, **k: None
    def step():
        r = m.train_step(b, 3, 0.5, 0.3); r["loss"].backward(); opt.step(max_grad_norm=5.0); opt.zero_grad()
    m.sync_prototypes = False
    for _ in range(4): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): step()
    m.param_sync(); e1.record(); torch.cuda.synchronize()
    if staged and not allreduce: dist.all_reduce = dist_all_reduce
    return e0.elapsed_time(e1) / K
print(f"one-call backward                      : {run(False, False):.3f} ms/step")
print(f"staged backward, all-reduce skipped    : {run(True, False):.3f} ms/step")
print(f"staged backward, 1-rank NCCL all-reduce: {run(True, True):.3f} ms/step")
print(f"one-call backward                      : {run(False, False):.3f} ms/step")
dist.destroy_process_group()
