"""Shared test helpers: build an oracle model + the CUDA model with identical weights, compare tensors."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import vlt5_oracle as O  # noqa: E402  (tests are the oracle's only importer besides smoke() and bench.py)


def oracle_config(layers=12, dec_layers=None, vocab=32200, feat_dim=2048, dropout=0.0, d_ff=3072):
    return O.VLT5Config(vocab_size=vocab, num_layers=layers, num_decoder_layers=layers if dec_layers is None else dec_layers,
                        feat_dim=feat_dim, dropout_rate=dropout, d_ff=d_ff)


def make_pair(layers=2, vocab=32200, feat_dim=2048, dropout=0.0, seed=66666, device="cuda", d_ff=3072, oracle_device=None):
    """(oracle fp32 model, vqacl_b200 model) with identical weights."""
    import vqacl_b200 as V
    ocfg = oracle_config(layers, None, vocab, feat_dim, dropout, d_ff)
    om = O.VLT5VQA(ocfg).init_weights_like_reference(seed)
    cfg = V.VLT5Config(vocab_size=vocab, num_layers=layers, num_decoder_layers=layers, feat_dim=feat_dim,
                       dropout_rate=dropout, d_ff=d_ff)
    m = V.VLT5VQA(cfg)
    res = m.load_state_dict(om.state_dict(), strict=True)
    m = m.to(device)
    om = om.to(oracle_device or device)
    return om, m


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-12)).item()


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    na, nb = a.norm().item(), b.norm().item()
    if na < 1e-20 and nb < 1e-20:
        return 1.0            # both exactly zero (e.g. attention over a single key has no q/k gradient)
    return (torch.dot(a, b) / max(na * nb, 1e-30)).item()
