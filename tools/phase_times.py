"""CUDA-event timing of the phases of one train step at the bench shape (B=320): where does the step go on the real
timeline (ncu's per-launch durations are serialised and over-count small launches)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch
import vqacl_b200 as V
import vlt5_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
torch.manual_seed(0)
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=drop)).to("cuda")
m.train()
opt = V.FusedAdamW(m)
eng = m._engine
batch = {k: v.cuda() for k, v in O.synthetic_batch(B, task_id=3).items()}
Ld, Le = 12, 12
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
acc = {}
for it in range(8):
    cb, keep, shape = m._stage_batch(batch["input_ids"], batch["vis_feats"], batch["boxes"], batch["target_ids"], batch["cate_labels"], batch["ques_labels"])
    eng.bind(*shape)
    marks = [("start", ev())]
    eng.forward_encoder(cb, it, True); marks.append(("enc fwd", ev()))
    ps = m._proto_state(True, 3, 0.5, 0.3)
    eng.forward_decoder(cb, ps, False, True); marks.append(("SI + dec fwd + head + CE", ev()))
    w_rows = torch.empty(B * 5, device="cuda")
    eng.loss_tail(keep[3], batch["scores"], B, 5, m._loss_buf, w_rows); marks.append(("loss tail", ev()))
    eng.backward(w_rows, False, 0, 1); marks.append(("bwd: CE + LM head", ev()))
    eng.backward(w_rows, False, 1, Ld + 1); marks.append(("bwd: decoder layers", ev()))
    eng.backward(w_rows, False, Ld + 1, Ld + 2); marks.append(("bwd: dec embed + cross-KV + enc final norm", ev()))
    eng.backward(w_rows, False, Ld + 2, Ld + 2 + Le); marks.append(("bwd: encoder layers", ev()))
    eng.backward(w_rows, False, Ld + 2 + Le, Ld + 3 + Le); marks.append(("bwd: embeddings", ev()))
    for p, gv in m._grad_views: p.grad = gv
    opt.step(max_grad_norm=5.0); marks.append(("clip + AdamW", ev()))
    torch.cuda.synchronize()
    if it >= 3:
        for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
        acc["total"] = acc.get("total", 0.0) + marks[0][1].elapsed_time(marks[-1][1])
for k, v in acc.items():
    print(f"{v / 5:8.3f} ms  {k}")
