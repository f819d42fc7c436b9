"""Generate tests/golden/memory_loss.pt by EXECUTING the reference's own `memory_loss` text (build container only).

The prototype pull loss is called at VL-T5/src/modeling_t5_our.py:591 but only DEFINED in the NExT-QA twin,
VL-T5/nextqa/modeling_t5_nextqa.py:544-555. Same `ast` lift as tools/gen_golden.py: the method runs
unmodified on an object that carries the two attributes it touches (Q_prototype, V_prototype). The fixture holds seeded
inputs, both losses and the autograd gradient of lambda_Q * loss_Q + lambda_V * loss_V (param.py:178-179 defaults; consumed
at vqacl.py:448-450) with respect to the encoder hidden states.

    python tools/gen_golden_memory_loss.py [/root/reference]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, lift  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "VL-T5", "nextqa", "modeling_t5_nextqa.py")


def main():
    ns = dict(torch=torch)
    exec(lift(SRC, "VLT5", only=["memory_loss"]), ns)
    m = ns["VLT5"].__new__(ns["VLT5"])
    g = torch.Generator().manual_seed(404)
    B, S = 6, 29                                   # 20 'question-side' + 9 'visual-side' rows (split hard-coded at 20, :380-381)
    m.Q_prototype = torch.randn(10, 768, generator=g)
    m.V_prototype = torch.randn(80, 768, generator=g)
    hidden = torch.randn(B, S, 768, generator=g).bfloat16().float().requires_grad_()   # bf16-representable: stored compactly
    ql = torch.zeros(B, 10)
    ql[torch.arange(B), torch.randint(0, 10, (B,), generator=g)] = 1
    cl = torch.zeros(B, 80)
    cl[torch.arange(B), torch.randint(0, 80, (B,), generator=g)] = 1
    lq, lv = m.memory_loss(hidden[:, :20, :], hidden[:, 20:, :], ql, cl)
    lam_q, lam_v = 0.01, 0.1
    (lam_q * lq + lam_v * lv).backward()
    os.makedirs(OUT, exist_ok=True)
    gh = hidden.grad
    # the gradient is constant over the tokens of each side (the loss sees only the token means): keep one row per side
    assert torch.equal(gh[:, :20], gh[:, :1].expand(-1, 20, -1)) and torch.equal(gh[:, 20:], gh[:, 20:21].expand(-1, S - 20, -1))
    torch.save(dict(hidden=hidden.detach().bfloat16(), ques_labels=ql, cate_labels=cl, Q_prototype=m.Q_prototype.clone(),
                    V_prototype=m.V_prototype.clone(), loss_Q=lq.detach().clone(), loss_V=lv.detach().clone(), lambda_Q=lam_q,
                    lambda_V=lam_v, grad_hidden_q_row=gh[:, 0].clone(), grad_hidden_v_row=gh[:, 20].clone()), os.path.join(OUT, "memory_loss.pt"))
    print("memory_loss.pt", os.path.getsize(os.path.join(OUT, "memory_loss.pt")), float(lq), float(lv))


if __name__ == "__main__":
    main()
