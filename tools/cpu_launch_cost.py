"""CPU cost of issuing one step's launches (tiny batch: the GPU is never the bottleneck)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch, vqacl_b200 as V, vlt5_oracle as O
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=0.1)).to("cuda"); m.train()
opt = V.FusedAdamW(m)
b = {k: v.cuda() for k, v in O.synthetic_batch(4, task_id=3).items()}
eng = m._engine
def step():
    r = m.train_step(b, 3, 0.5, 0.3); r["loss"].backward(); opt.step(max_grad_norm=5.0); opt.zero_grad()
for _ in range(3): step()
torch.cuda.synchronize()
l0 = eng.launch_count(); t0 = time.perf_counter()
for _ in range(3):
    step(); torch.cuda.synchronize()
t1 = time.perf_counter()
n = eng.launch_count() - l0
print(f"{n/3:.0f} launches/step, wall {1e3*(t1-t0)/3:.2f} ms/step incl. sync -> {1e6*(t1-t0)/n:.1f} us per launch upper bound")
# pure issue time of forward_encoder
cb, keep, shape = m._stage_batch(b["input_ids"], b["vis_feats"], b["boxes"], b["target_ids"], b["cate_labels"], b["ques_labels"])
eng.bind(*shape)
torch.cuda.synchronize()
l0 = eng.launch_count(); t0 = time.perf_counter()
eng.forward_encoder(cb, 1, True)
t1 = time.perf_counter()
print(f"forward_encoder: {eng.launch_count()-l0} launches issued in {1e6*(t1-t0):.0f} us -> {1e6*(t1-t0)/(eng.launch_count()-l0):.2f} us/launch")
torch.cuda.synchronize()
