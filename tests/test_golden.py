"""Pin the oracle against outputs of the reference's own source (tests/golden, made by tools/gen_golden.py)."""
import os

import pytest
import torch

from helpers import O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_visual_embedding_matches_reference():
    d = torch.load(os.path.join(G, "visual_embedding.pt"))
    cfg = O.VLT5Config(vocab_size=d["vocab"], feat_dim=d["feat_dim"])
    shared = torch.nn.Embedding(d["vocab"], 768)
    ve = O.VisualEmbedding(cfg, shared)
    ve.load_state_dict(d["state"])
    with torch.no_grad():
        out = ve(d["feats"], d["boxes"])
    torch.testing.assert_close(out, d["out"], rtol=1e-6, atol=1e-6)


def test_prototype_state_machine_matches_reference():
    d = torch.load(os.path.join(G, "prototype_path.pt"))
    bank = O.PrototypeBank()
    for st in d["steps"]:
        h = st["hidden"].float()
        curQ, numQ = O.calculate_current_prototype(h[:, :20], st["ques_labels"])
        curV, numV = O.calculate_current_prototype(h[:, 20:], st["cate_labels"])
        torch.testing.assert_close(curQ, st["curQ"], rtol=0, atol=0)
        torch.testing.assert_close(curV, st["curV"], rtol=0, atol=0)
        assert torch.equal(numQ, st["numQ"]) and torch.equal(numV, st["numV"])
        bank.update(curQ, curV, numQ, numV, st["task"], d["alpha"], d["beta"])
        torch.testing.assert_close(bank.Q_prototype, st["Q_prototype"], rtol=0, atol=0)
        torch.testing.assert_close(bank.V_prototype, st["V_prototype"], rtol=0, atol=0)
        assert torch.equal(bank.Q_prototype_num, st["Q_num"]) and torch.equal(bank.V_prototype_num, st["V_num"])
        rq, iq = O.cosine_similarity_multi(bank.Q_prototype, h[:, :20].mean(1))
        rv, iv = O.cosine_similarity_multi(bank.V_prototype, h[:, 20:].mean(1))
        assert torch.equal(iq, st["idx_Q"]) and torch.equal(iv, st["idx_V"])
        torch.testing.assert_close(rq, st["retr_Q"], rtol=0, atol=0)
        torch.testing.assert_close(rv, st["retr_V"], rtol=0, atol=0)
    r, i = O.cosine_similarity_multi(d["eval_P"], d["eval_x"])
    assert torch.equal(i, d["eval_idx"])
    torch.testing.assert_close(r, d["eval_retr"], rtol=0, atol=0)


def test_loss_tail_matches_reference():
    d = torch.load(os.path.join(G, "loss_tail.pt"))
    labels, rows, scores = d["labels"], d["loss_rows"], d["scores"]
    B, T = labels.shape
    m = (labels != -100).float()
    loss = ((rows.view(B, T) * m).sum(1) / m.sum(1).clamp(min=1) * scores).mean()
    torch.testing.assert_close(loss, d["loss"], rtol=0, atol=0)
