"""GPU bring-up: one VQACL train step of the CUDA engine vs the fp32 oracle on identical weights / inputs.
python tools/step_check.py [layers] [B] [vocab]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import O, make_pair, rel_err, cos

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
vocab = int(sys.argv[3]) if len(sys.argv) > 3 else 32200
torch.manual_seed(0)
om, m = make_pair(layers=layers, vocab=vocab)
om.train(); m.train()
task = 3
# put a non-trivial inherited bank in both
g = torch.Generator().manual_seed(7)
Q0 = torch.randn(10, 768, generator=g); V0 = torch.randn(80, 768, generator=g)
om.bank.Q_prototype = Q0.clone().cuda(); om.bank.V_prototype = V0.clone().cuda()
m.Q_prototype = Q0; m.V_prototype = V0
import vqacl_b200 as V
opt = V.FusedAdamW(m, lr=1e-4, eps=1e-6, weight_decay=0.01)
oopt = O.HFAdamW(list(om.named_parameters()))
for step in range(3):
    batch = O.synthetic_batch(B, seed=1234 + step, task_id=task, vocab=min(32000, vocab))
    ro = om.train_step(batch, task, 0.5, 0.3)
    r = m.train_step(batch, task, 0.5, 0.3)
    torch.cuda.synchronize()
    print(f"step {step}: loss ours {r['loss'].item():.6f} oracle {ro['loss'].item():.6f} rel {abs(r['loss'].item()-ro['loss'].item())/abs(ro['loss'].item()):.3e}")
    print("  enc hidden rel", rel_err(r["encoder_hidden_states"], ro["encoder_hidden_states"]))
    print("  logits rel", rel_err(r["logits"], ro["logits"]), "cos", cos(r["logits"], ro["logits"]))
    print("  idxQ equal", torch.equal(r["max_idx_Q"], ro["max_idx_Q"]), "idxV equal", torch.equal(r["max_idx_V"], ro["max_idx_V"]),
          (r["max_idx_V"] != ro["max_idx_V"]).sum().item())
    print("  Qproto rel", rel_err(m.Q_prototype, om.bank.Q_prototype), "Vproto rel", rel_err(m.V_prototype, om.bank.V_prototype))
    print("  counts equal", torch.equal(m.Q_prototype_num, om.bank.Q_prototype_num), torch.equal(m.V_prototype_num, om.bank.V_prototype_num))
    r["loss"].backward()
    ro["loss"].backward()
    torch.cuda.synchronize()
    onamed = dict(om.named_parameters())
    worst = []
    for n, p in m.named_parameters():
        if p.grad is None:
            assert onamed[n].grad is None, n
            continue
        worst.append((cos(p.grad, onamed[n].grad), rel_err(p.grad, onamed[n].grad), n))
    worst.sort()
    for c, e, n in worst[:6]:
        print(f"  grad {n}: cos {c:.5f} rel {e:.3e}")
    print("  median cos", worst[len(worst) // 2][0])
    gn = torch.nn.utils.clip_grad_norm_([p for p in om.parameters() if p.grad is not None], 5.0)
    oopt.step()
    for p in om.parameters():
        p.grad = None
    opt.step(max_grad_norm=5.0)
    opt.zero_grad()
    torch.cuda.synchronize()
    print("  grad norm ours", opt.grad_sumsq.sqrt().item(), "oracle", gn.item())
    pw = max(rel_err(p, onamed[n]) for n, p in m.named_parameters())
    print("  params after step: worst rel", pw)
print("launches", m._engine.launch_count())
