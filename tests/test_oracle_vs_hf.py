"""Cross-check the oracle's T5 leaf maths against the installed transformers (5.5) modules whose arithmetic is shared
with the 4.2.1 release the reference imports. Skipped if transformers is unavailable."""
import pytest
import torch

from helpers import O

hf = pytest.importorskip("transformers.models.t5.modeling_t5")
from transformers import T5Config  # noqa: E402


def hf_cfg(layers=2):
    return T5Config(vocab_size=512, d_model=768, d_kv=64, d_ff=3072, num_layers=layers, num_decoder_layers=layers, num_heads=12,
                    relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                    feed_forward_proj="relu", tie_word_embeddings=True, decoder_start_token_id=0, pad_token_id=0, eos_token_id=1)


def test_layernorm():
    torch.manual_seed(0)
    x = torch.randn(5, 7, 768) * 3
    a, b = O.T5LayerNorm(768), hf.T5LayerNorm(768)
    w = torch.randn(768)
    a.weight.data.copy_(w); b.weight.data.copy_(w)
    torch.testing.assert_close(a(x), b(x), rtol=0, atol=0)


@pytest.mark.parametrize("bidirectional", [True, False])
def test_relative_position_bucket(bidirectional):
    rel = torch.arange(-130, 131)[None, :] - torch.zeros(1, 1, dtype=torch.long)
    ours = O.relative_position_bucket(rel, bidirectional, 32, 128)
    theirs = hf.T5Attention._relative_position_bucket(rel, bidirectional=bidirectional, num_buckets=32, max_distance=128)
    assert torch.equal(ours, theirs)
    # the host table the CUDA kernels consume
    from vqacl_b200.engine import rel_bucket_table
    tab = rel_bucket_table(bidirectional, 32, 128)
    assert torch.equal(tab.long(), theirs[0, 130 - 63:130 + 64])


def test_dense_relu_dense_and_ff_layer():
    torch.manual_seed(1)
    cfg, ocfg = hf_cfg(), O.VLT5Config(vocab_size=512, dropout_rate=0.0)
    a, b = O.T5LayerFF(ocfg).eval(), hf.T5LayerFF(cfg).eval()
    b.DenseReluDense.wi.weight.data.copy_(a.DenseReluDense.wi.weight)
    b.DenseReluDense.wo.weight.data.copy_(a.DenseReluDense.wo.weight)
    b.layer_norm.weight.data.copy_(a.layer_norm.weight)
    x = torch.randn(2, 9, 768)
    torch.testing.assert_close(a(x), b(x), rtol=1e-5, atol=1e-5)


def test_text_only_seq2seq_matches_hf():
    """Whole text-only encoder/decoder stack + tied LM head + CE vs T5ForConditionalGeneration with copied weights.
    (Mask constant differs: 5.5 uses finfo.min, 4.2.1 uses -10000/-1e9; with no padding both are inactive.)"""
    torch.manual_seed(2)
    cfg = hf_cfg(2)
    ref = hf.T5ForConditionalGeneration(cfg).eval()
    ocfg = O.VLT5Config(vocab_size=512, num_layers=2, num_decoder_layers=2, dropout_rate=0.0)
    om = O.VLT5VQA(ocfg).eval()
    sd = ref.state_dict()
    mine = om.state_dict()
    for k in mine:
        if k in sd and mine[k].shape == sd[k].shape:
            mine[k].copy_(sd[k])
    om.shared.weight.data.copy_(sd["shared.weight"])
    ids = torch.randint(2, 512, (3, 11))
    labels = torch.randint(2, 512, (3, 6))
    with torch.no_grad():
        out = ref(input_ids=ids, labels=labels)
        # oracle pieces, text only (no visual tokens, no prototype rows)
        x = om.shared(ids)
        bias = om.encoder.block[0].layer[0].SelfAttention.compute_bias(11, 11)
        for blk in om.encoder.block:
            x = blk(x, bias)
        enc = om.encoder.final_layer_norm(x)
        torch.testing.assert_close(enc, out.encoder_last_hidden_state, rtol=2e-4, atol=2e-4)
        logits, _ = om.decode_logits(om.shift_right(labels), enc, ids)
        torch.testing.assert_close(logits, out.logits, rtol=2e-4, atol=2e-4)


def test_init_std_table():
    """HF `_init_weights` scheme as applied by the oracle and by vqacl_b200.VLT5.init_weights."""
    import vqacl_b200 as V
    torch.manual_seed(3)
    m = V.VLT5VQA(V.VLT5Config(num_layers=1, num_decoder_layers=1, vocab_size=2048))
    att = m.encoder.block[0].layer[0].SelfAttention
    assert abs(att.q.weight.std().item() - (768 * 64) ** -0.5) < 2e-4
    assert abs(att.k.weight.std().item() - 768 ** -0.5) < 2e-3
    assert abs(m.encoder.block[0].layer[1].DenseReluDense.wo.weight.std().item() - 3072 ** -0.5) < 1e-3
    assert abs(m.shared.weight.std().item() - 1.0) < 2e-2
    assert m.lm_head.weight is m.shared.weight
    assert m.encoder.visual_embedding.obj_order_embedding.weight is m.shared.weight


def test_greedy_loop_matches_hf_generate():
    """The oracle's greedy loop (start token, argmax, pad after EOS, stop when every row is finished or at max_length;
    SURVEY.md §8 a16) against transformers' own `generate` on the same decoder, memory and cross mask. A tiny vocabulary
    makes rows hit EOS at different steps, so the finished-row bookkeeping is exercised."""
    from transformers.modeling_outputs import BaseModelOutput
    torch.manual_seed(2)
    V = 8
    cfg = T5Config(vocab_size=V, d_model=768, d_kv=64, d_ff=128, num_layers=2, num_decoder_layers=2, num_heads=12,
                   relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                   feed_forward_proj="relu", tie_word_embeddings=True, decoder_start_token_id=0, pad_token_id=0, eos_token_id=1)
    ref = hf.T5ForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        ref.shared.weight.mul_(0.02)   # small tied embeddings: the next token is not simply the current one again
    om = O.VLT5VQA(O.VLT5Config(vocab_size=V, d_ff=128, num_layers=2, num_decoder_layers=2, dropout_rate=0.0)).eval()
    sd, mine = ref.state_dict(), om.state_dict()
    for k in mine:
        if k in sd and mine[k].shape == sd[k].shape:
            mine[k].copy_(sd[k])
    om.shared.weight.data.copy_(sd["shared.weight"])
    B, L, S2 = 12, 9, 14
    ids = torch.randint(2, V, (B, L))
    ids[3, 6:] = 0                                     # padded text positions are masked in cross-attention
    ids[7, 4:] = 0
    mem = torch.randn(B, S2, 768)
    mask = torch.cat([ids.ne(0).long(), torch.ones(B, S2 - L, dtype=torch.long)], 1)
    mine_tok = om.greedy_decode(mem, ids, max_length=20)
    with torch.no_grad():
        theirs = ref.generate(encoder_outputs=BaseModelOutput(last_hidden_state=mem), attention_mask=mask, max_length=20,
                              do_sample=False, num_beams=1)
    n = min(mine_tok.size(1), theirs.size(1))
    assert torch.equal(mine_tok[:, :n], theirs[:, :n])
    assert (mine_tok[:, n:] == 0).all() and (theirs[:, n:] == 0).all()
    finished_at = [(row == 1).nonzero()[0].item() if (row == 1).any() else -1 for row in mine_tok]
    assert len({f for f in finished_at}) >= 3          # rows really stop at different steps (else the test is vacuous)


def test_adamw_core_matches_torch_when_eps_and_decay_vanish():
    """transformers 4.2.1's AdamW is not installed anywhere (5.5 removed it), so its two distinguishing choices — eps added
    to sqrt(v) BEFORE the bias correction, weight decay applied AFTER the Adam step with lr — stay "parity unpinned"
    (DESIGN.md §6). Everything else (moment updates, both bias corrections, step size) coincides with torch.optim.AdamW
    once eps -> 0 and weight_decay = 0, which is what this test pins."""
    torch.manual_seed(9)
    w0 = torch.randn(37, 5)
    a = torch.nn.Parameter(w0.clone()); b = torch.nn.Parameter(w0.clone())
    mine = O.HFAdamW([("w", a)], lr=1e-2, betas=(0.9, 0.999), eps=1e-30, weight_decay=0.0)
    ref = torch.optim.AdamW([b], lr=1e-2, betas=(0.9, 0.999), eps=1e-30, weight_decay=0.0)
    for _ in range(7):
        g = torch.randn(37, 5)
        a.grad = g.clone(); b.grad = g.clone()
        mine.step(); ref.step()
    torch.testing.assert_close(a.data, b.data, rtol=2e-6, atol=2e-7)
    # and the decay it applies is p <- p - lr * wd * p on the already-stepped parameter
    c = torch.nn.Parameter(w0.clone())
    dec = O.HFAdamW([("w", c)], lr=1e-2, eps=1e-6, weight_decay=0.1)
    nodec = O.HFAdamW([("w", a)], lr=1e-2, eps=1e-6, weight_decay=0.0)
    a.data.copy_(w0); g = torch.randn(37, 5); a.grad = g.clone(); c.grad = g.clone()
    nodec.step(); dec.step()
    torch.testing.assert_close(c.data, a.data * (1 - 1e-2 * 0.1), rtol=1e-6, atol=1e-7)
