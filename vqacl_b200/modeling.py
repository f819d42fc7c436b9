"""Drop-in model classes for the VQACL hot path, mirroring the reference's Python model API.

    reference                                           here
    VLT5VQA            VL-T5/src/vqa_model.py:10-121    VLT5VQA
    VLT5               VL-T5/src/modeling_t5_our.py:342-772   VLT5
    JointEncoder       :145-339                          parameter container `encoder` (+ `encode()`)
    VisualEmbedding    :27-143                           parameter container `encoder.visual_embedding`
    VLSeq2SeqLMOutput  :774-835                          VLSeq2SeqLMOutput

Same constructor / `from_pretrained` / `resize_token_embeddings` / `train_step` / `test_step` / `forward` / `generate`
signatures, same `state_dict` keys, same `Q_prototype` / `V_prototype` attributes. The arithmetic is NOT here: once the
module is moved to a CUDA device every parameter becomes a view into the native engine's flat fp32 arena and every
kernel of forward, backward, optimizer and greedy decode is launched from C++ (vqacl_b200/csrc). On a CPU device the
module is only a parameter container (init, load_state_dict, resize) — calling the hot path there raises: there is no
CPU fallback.
"""
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Optional

import torch
import torch.nn as nn

from ._lib import VqaclError
from .config import VLT5Config
from .engine import CProtoState, Engine


# ------------------------------------------------------------------------------------------------------------------
# parameter containers (names = the reference's state_dict keys)
# ------------------------------------------------------------------------------------------------------------------
class T5LayerNorm(nn.Module):
    """Weight holder of HF T5LayerNorm (RMS norm; computed by rmsnorm_fwd/bwd kernels)."""

    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _Attention(nn.Module):
    def __init__(self, cfg, has_relative_attention_bias):
        super().__init__()
        inner = cfg.num_heads * cfg.d_kv
        self.has_relative_attention_bias = has_relative_attention_bias
        self.q = nn.Linear(cfg.d_model, inner, bias=False)
        self.k = nn.Linear(cfg.d_model, inner, bias=False)
        self.v = nn.Linear(cfg.d_model, inner, bias=False)
        self.o = nn.Linear(inner, cfg.d_model, bias=False)
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(cfg.relative_attention_num_buckets, cfg.num_heads)


class _SelfAttnLayer(nn.Module):
    def __init__(self, cfg, has_bias):
        super().__init__()
        self.SelfAttention = _Attention(cfg, has_bias)
        self.layer_norm = T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon)


class _CrossAttnLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.EncDecAttention = _Attention(cfg, False)
        self.layer_norm = T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon)


class _DenseReluDense(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.wi = nn.Linear(cfg.d_model, cfg.d_ff, bias=False)
        self.wo = nn.Linear(cfg.d_ff, cfg.d_model, bias=False)


class _FFLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.DenseReluDense = _DenseReluDense(cfg)
        self.layer_norm = T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon)


class _Block(nn.Module):
    def __init__(self, cfg, is_decoder, has_bias):
        super().__init__()
        self.layer = nn.ModuleList([_SelfAttnLayer(cfg, has_bias)])
        if is_decoder:
            self.layer.append(_CrossAttnLayer(cfg))
        self.layer.append(_FFLayer(cfg))


class VisualEmbedding(nn.Module):
    """Parameters of modeling_t5_our.py:27-76 (flags at their defaults; config.check_supported enforces it)."""

    def __init__(self, cfg, obj_order_embedding):
        super().__init__()
        self.feat_embedding = nn.Sequential(nn.Linear(cfg.feat_dim, cfg.d_model), T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon))
        self.absolute_vis_pos_embedding = nn.Sequential(nn.Linear(cfg.pos_dim + 1, cfg.d_model),
                                                        T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon))
        self.obj_order_embedding = obj_order_embedding
        self.img_order_embedding = nn.Embedding(cfg.n_images, cfg.d_model)


class JointEncoder(nn.Module):
    def __init__(self, cfg, embed_tokens):
        super().__init__()
        self.embed_tokens = embed_tokens
        self.visual_embedding = VisualEmbedding(cfg, embed_tokens)
        self.block = nn.ModuleList([_Block(cfg, False, i == 0) for i in range(cfg.num_layers)])
        self.final_layer_norm = T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon)


class _DecoderStack(nn.Module):
    def __init__(self, cfg, embed_tokens):
        super().__init__()
        self.embed_tokens = embed_tokens
        self.block = nn.ModuleList([_Block(cfg, True, i == 0) for i in range(cfg.num_decoder_layers)])
        self.final_layer_norm = T5LayerNorm(cfg.d_model, cfg.layer_norm_epsilon)


@dataclass
class VLSeq2SeqLMOutput:
    """Field-compatible with modeling_t5_our.py:774-835 (dict-style and attribute access)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Any = None
    decoder_last_hidden_state: Any = None
    decoder_hidden_states: Any = None
    encoder_hidden_states: Optional[torch.Tensor] = None
    encoder_attention_mask: Optional[torch.Tensor] = None
    loss_memory_Q: Any = 0
    loss_memory_V: Any = 0
    max_idx_Q: Optional[torch.Tensor] = None
    max_idx_V: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        return getattr(self, k)

    def __contains__(self, k):
        return getattr(self, k, None) is not None


# ------------------------------------------------------------------------------------------------------------------
# autograd bridges: loss.backward() (vqacl.py:461) runs the native backward and leaves param.grad as arena views
# ------------------------------------------------------------------------------------------------------------------
class _StepLoss(torch.autograd.Function):
    """Scalar loss of train_step; backward(g) = native backward with dL/dloss_row = g * w_rows."""

    @staticmethod
    def forward(ctx, anchor, model, loss_buf, w_rows):
        ctx.model, ctx.w_rows = model, w_rows
        return loss_buf.reshape(()).clone()

    @staticmethod
    def backward(ctx, g):
        ctx.model._native_backward(ctx.w_rows, g)
        return None, None, None, None


class _StepLossMem(torch.autograd.Function):
    """train_step / forward with memory=True: (main loss or loss rows, loss_memory_Q, loss_memory_V) from ONE node, so that the
    single native backward sees all three upstream gradients at once (the caller combines them as
    loss + lambda_Q * loss_memory_Q + lambda_V * loss_memory_V, vqacl.py:448-450)."""

    @staticmethod
    def forward(ctx, anchor, model, main, w_rows, mem2):
        ctx.model, ctx.w_rows = model, w_rows
        return main.clone() if w_rows is None else main.reshape(()).clone(), mem2[0].clone(), mem2[1].clone()

    @staticmethod
    def backward(ctx, g, gq, gv):
        m = ctx.model
        m._mem_g = torch.stack([gq.reshape(()), gv.reshape(())]).to(torch.float32).contiguous()
        m._engine.set_memory_loss_grads(m._mem_g)
        try:
            if ctx.w_rows is None:
                m._native_backward(g.contiguous().float(), None)       # row losses: g is dL/dloss_row
            else:
                m._native_backward(ctx.w_rows, g)
        finally:
            m._engine.set_memory_loss_grads(None)
        return None, None, None, None, None


class _RowLoss(torch.autograd.Function):
    """Per-row CE loss of VLT5.forward (reduction='none'); backward(g[B*T]) = native backward with w_rows = g."""

    @staticmethod
    def forward(ctx, anchor, model, loss_rows):
        ctx.model = model
        return loss_rows.clone()

    @staticmethod
    def backward(ctx, g):
        ctx.model._native_backward(g.contiguous().float(), None)
        return None, None, None


def plan_grad_buckets(stage_ranges, min_elems):
    """Which arena ranges to all-reduce after which backward stage.

    `stage_ranges[s] = (a, b)` is the gradient-arena range that is final once stage s has run (possibly empty). Adjacent
    ranges are merged until a bucket holds at least `min_elems` elements (NVSwitch all-reduce cost is launch-latency bound for
    small messages, not link bound); a non-adjacent range or the end of backward flushes what is pending.
    Returns {stage: [(a, b), ...]}; every element of every non-empty range is covered exactly once."""
    out = {}
    pend = None
    last = len(stage_ranges) - 1
    for s, (a, b) in enumerate(stage_ranges):
        if b > a:
            if pend is None:
                pend = [a, b]
            elif a == pend[1]:
                pend[1] = b
            elif b == pend[0]:
                pend[0] = a
            else:
                out.setdefault(s, []).append(tuple(pend))
                pend = [a, b]
        if pend is not None and (pend[1] - pend[0] >= min_elems or s == last):
            out.setdefault(s, []).append(tuple(pend))
            pend = None
    return out


def plan_shards(stage_ranges, n_tail, n_train, min_elems, world):
    """Reduce-scatter plan of the sharded step tail (N > 1 GPUs).

    The arena is [0, n_tail) GEMM-only matrices | [n_tail, n_train) parameters every rank reads in fp32 (engine layout).
    The first region is reduce-scattered in buckets (the stage ranges clipped to it, merged like plan_grad_buckets): rank r
    then owns elements [a + r * (b - a) / world, a + (r + 1) * (b - a) / world) of bucket (a, b), runs AdamW on them and
    all-gathers the refreshed bf16 weights into the same bucket. The tail is all-reduced after the last stage and updated by
    every rank. Returns (flush_after {stage: [(a, b), ...]}, buckets [(a, b), ...] in backward order, tail (a, b))."""
    clipped = [(min(a, n_tail), min(b, n_tail)) for a, b in stage_ranges]
    flush_after = plan_grad_buckets(clipped, min_elems)
    buckets = [ab for s in sorted(flush_after) for ab in flush_after[s]]
    for a, b in buckets:
        if (b - a) % (8 * world) or a % 8:
            raise VqaclError(f"bucket [{a}, {b}) cannot be split into {world} 32-byte aligned slices")
    return flush_after, buckets, (n_tail, n_train)


def owned_slices(buckets, world, rank):
    return [(a + rank * ((b - a) // world), a + (rank + 1) * ((b - a) // world)) for a, b in buckets]


def chunk_events(chunks, buckets_in_gather_order):
    """For every parameter chunk (a, b) the index of the LAST gathered bucket that overlaps it (None: no sharded data in it)."""
    out = []
    for ca, cb in chunks:
        last = None
        for i, (a, b) in enumerate(buckets_in_gather_order):
            if a < cb and ca < b:
                last = i
        out.append(last)
    return out


class VLT5(nn.Module):
    def __init__(self, config: VLT5Config):
        super().__init__()
        self.config = config
        self.model_dim = config.d_model
        self.shared = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = JointEncoder(config, self.shared)
        self.decoder = _DecoderStack(config, self.shared)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.prototype_fc1 = nn.Linear(config.d_model, config.d_model)   # modeling_t5_our.py:379-380 (never used)
        self.prototype_fc2 = nn.Linear(config.d_model, config.d_model)
        self.L = config.proto_split_L        # :381
        self.V_L = 36                        # :382
        self.init_weights()
        self.model_parallel = False
        self.device_map = None
        # SI prototype bank bookkeeping (:391-396). The banks themselves are device buffers created on first use.
        self.Q_task_mem_proto = {}
        self.V_task_mem_proto = {}
        self.Q_task_cur_proto = {}
        self.V_task_cur_proto = {}
        self._Q_prototype = None
        self._V_prototype = None
        self._Q_prototype_num = None
        self._V_prototype_num = None
        # native engine state
        self._engine = None
        self._anchor = None
        self._step_seed = 0
        self._base_seed = 0x5EED
        self.sync_prototypes = True      # multi-GPU: all-reduce class sums so every rank holds the global-batch bank
        self.sync_grads = True           # multi-GPU: all-reduce(avg) gradients (what the reference's DDP wrap intends)
        self._comm_stream = None
        self.grad_bucket_elems = 8 << 20   # ~32 MB fp32 per NCCL all-reduce bucket
        self.comm_sms = 0                  # SMs to leave to NCCL during backward (0: none reserved — measured neutral at 8 GPUs)
        # N > 1: reduce-scatter gradients instead of all-reducing them and let FusedAdamW update only this rank's slices
        # (set by FusedAdamW(shard_state=...)); fp32 masters of the other ranks' slices are refreshed by gather_params()
        self.shard_optimizer = False
        self._shard_plan = None
        self._masters_stale = False
        self._param_events = None

    # -- construction helpers the reference calls --------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, name="t5-base", config=None, **kwargs):
        """The reference loads hub weights and (with --from_scratch, every script) re-randomises them
        (trainer_base.py:218-238). There is no hub here: this returns a randomly initialised model; real weights come
        through load_state_dict."""
        if config is None:
            config = VLT5Config.from_pretrained(name)
        return cls(config, **kwargs)

    def tie_weights(self):
        self.lm_head.weight = self.shared.weight
        self.encoder.embed_tokens = self.shared
        self.decoder.embed_tokens = self.shared
        self.encoder.visual_embedding.obj_order_embedding = self.shared

    @torch.no_grad()
    def init_weights(self):
        """HF T5 `_init_weights` (hf5.5 modeling_t5.py:541-593) + weight tying. Layers HF's scheme does
        not know (visual Linear layers, img_order_embedding, prototype_fc*) keep whatever they hold (H14)."""
        c = self.config
        f, d, dk, H, ff = c.initializer_factor, c.d_model, c.d_kv, c.num_heads, c.d_ff
        for m in self.modules():
            if isinstance(m, T5LayerNorm):
                m.weight.fill_(f * 1.0)
            elif isinstance(m, _DenseReluDense):
                m.wi.weight.normal_(0.0, f * d ** -0.5)
                m.wo.weight.normal_(0.0, f * ff ** -0.5)
            elif isinstance(m, _Attention):
                m.q.weight.normal_(0.0, f * (d * dk) ** -0.5)
                m.k.weight.normal_(0.0, f * d ** -0.5)
                m.v.weight.normal_(0.0, f * d ** -0.5)
                m.o.weight.normal_(0.0, f * (H * dk) ** -0.5)
                if m.has_relative_attention_bias:
                    m.relative_attention_bias.weight.normal_(0.0, f * d ** -0.5)
        self.shared.weight.normal_(0.0, f * 1.0)
        self.tie_weights()
        self._mark_params_dirty()

    def resize_token_embeddings(self, new_num_tokens):
        """vqacl.py:98-99: 32128 -> 32200 (tokenization.py:58-60). Only before the model is moved to the GPU."""
        if self._engine is not None:
            raise VqaclError("resize_token_embeddings must be called before .to(cuda) (the arena layout is fixed then)")
        old = self.shared.weight.data
        n = min(old.size(0), new_num_tokens)
        new = nn.Embedding(new_num_tokens, self.config.d_model)
        new.weight.data.normal_(0.0, self.config.initializer_factor * 1.0)
        new.weight.data[:n] = old[:n]
        self.shared = new
        self.lm_head = nn.Linear(self.config.d_model, new_num_tokens, bias=False)
        self.config.vocab_size = new_num_tokens
        self.tie_weights()
        return self.shared

    def get_input_embeddings(self):
        return self.shared

    # -- device placement: pack parameters into the engine arena -------------------------------------------------------
    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, dtype=torch.float32, device=self.shared.weight.device))
        if probe.dtype != torch.float32:
            raise VqaclError("master weights are fp32 (the kernels keep their own bf16 copies); .half()/.bfloat16() is not supported")
        if probe.device.type == "cuda":
            if self._engine is None:
                self._pack(probe.device)
            elif probe.device != self._engine.device:
                raise VqaclError("moving a packed model between CUDA devices is not supported")
            return self
        if self._engine is not None:
            raise VqaclError("a packed model stays on its CUDA device; use state_dict() to take weights to the CPU")
        return super()._apply(fn, recurse)

    def _named_unique_params(self):
        seen = {}
        for name, p in self.named_parameters(remove_duplicate=False):
            seen.setdefault(id(p), (name, p))
        return list(seen.values())

    @torch.no_grad()
    def _pack(self, device):
        self.config.check_supported()
        eng = Engine(self.config, device)
        alias = {"encoder.embed_tokens.weight": "shared.weight", "decoder.embed_tokens.weight": "shared.weight",
                 "lm_head.weight": "shared.weight", "encoder.visual_embedding.obj_order_embedding.weight": "shared.weight"}
        self._grad_views = []
        for name, p in self._named_unique_params():
            key = alias.get(name, name)
            if key not in eng.table:
                raise VqaclError(f"parameter {name} has no slot in the engine arena")
            off, rows, cols, grp = eng.table[key]
            if rows * cols != p.numel():
                raise VqaclError(f"parameter {name}: {tuple(p.shape)} does not match the arena slot {rows}x{cols}")
            view = eng.P[off:off + p.numel()].view(p.shape)
            view.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = view
            p.grad = None
            if grp != 2:
                self._grad_views.append((p, eng.G[off:off + p.numel()].view(p.shape)))
        self._engine = eng
        self._anchor = torch.zeros((), device=device, requires_grad=True)
        self._loss_buf = torch.zeros(4, dtype=torch.float32, device=device)
        eng.bf16_stale = True
        d = self.config.d_model
        for attr, n in (("_Q_prototype", self.config.n_ques_classes), ("_V_prototype", self.config.n_cate_classes)):
            cur = getattr(self, attr)
            buf = torch.zeros(n, d, dtype=torch.float32, device=device)
            if cur is not None:
                buf.copy_(cur)
            setattr(self, attr, buf)
        for attr, n in (("_Q_prototype_num", self.config.n_ques_classes), ("_V_prototype_num", self.config.n_cate_classes)):
            cur = getattr(self, attr)
            buf = torch.zeros(n, dtype=torch.float32, device=device)
            if cur is not None:
                buf.copy_(cur)          # counts restored from a checkpoint before the model was moved to the GPU
            setattr(self, attr, buf)

    def param_sync(self):
        """Make the current stream wait for a pending overlapped optimizer step (FusedAdamW(overlap_with_next_forward=True))."""
        if getattr(self, "_engine", None) is not None:
            self._engine.param_sync()

    def grad_shard_plan(self):
        """(flush_after, buckets, tail) of the sharded step tail for the current world size (cached)."""
        world = self._world()
        if self._shard_plan is None or self._shard_plan[0] != world:
            eng = self._engine
            ranges = [eng.backward_stage_range(s) for s in range(eng.n_backward_stages())]
            self._shard_plan = (world,) + plan_shards(ranges, eng.n_tail, eng.n_train, self.grad_bucket_elems, world)
        return self._shard_plan[1:]

    def gather_params(self):
        """Sharded optimizer: all-gather the fp32 master weights of the GEMM matrices (each rank only updates its own slices;
        the forward runs on the all-gathered bf16 copies). Needed before reading parameters: state_dict() calls it."""
        if not self._masters_stale:
            return
        import torch.distributed as dist
        eng = self._engine
        self.param_sync()
        _, buckets, _ = self.grad_shard_plan()
        world, rank = dist.get_world_size(), dist.get_rank()
        for (a, b), (oa, ob) in zip(buckets, owned_slices(buckets, world, rank)):
            dist.all_gather_into_tensor(eng.P[a:b], eng.P[oa:ob])
        self._masters_stale = False

    def state_dict(self, *a, **kw):
        self.param_sync()
        self.gather_params()
        if getattr(self, "_engine", None) is not None:
            self._engine.check_device_errors()      # checkpoint time at the latest (the step itself never synchronises)
        return super().state_dict(*a, **kw)

    def _mark_params_dirty(self):
        """fp32 masters were modified outside the fused optimizer -> bf16 GEMM copies must be refreshed before use."""
        if getattr(self, "_engine", None) is not None:
            self._engine.bf16_stale = True

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in state_dict.items())
        self.param_sync()       # a pending overlapped optimizer step must not overwrite the weights loaded below
        res = super().load_state_dict(sd, strict=strict, **kw)
        self.tie_weights()
        self._mark_params_dirty()
        self._masters_stale = False       # every fp32 master was just overwritten on every rank
        return res

    def apply(self, fn):
        r = super().apply(fn)
        self._mark_params_dirty()
        return r

    def _need_engine(self):
        if self._engine is None:
            raise VqaclError("the VQACL hot path runs on CUDA only (hand-written sm_100a kernels; no CPU fallback): "
                             "move the model with .to('cuda') first")
        return self._engine

    # -- SI prototype banks as plain attributes (vqacl.py:420-423, 541-542) --------------------------------------------
    @property
    def Q_prototype(self):
        return self._Q_prototype

    @Q_prototype.setter
    def Q_prototype(self, t):
        self._set_bank("_Q_prototype", t)

    @property
    def V_prototype(self):
        return self._V_prototype

    @V_prototype.setter
    def V_prototype(self, t):
        self._set_bank("_V_prototype", t)

    @property
    def Q_prototype_num(self):
        return self._Q_prototype_num

    @property
    def V_prototype_num(self):
        return self._V_prototype_num

    def _set_bank(self, attr, t):
        cur = getattr(self, attr)
        if cur is not None and cur.is_cuda:
            cur.copy_(t.detach().to(cur.device, torch.float32))
        else:
            setattr(self, attr, t.detach().float().clone())

    # -- hot path ------------------------------------------------------------------------------------------------------
    def _stage_batch(self, input_ids, vis_feats, boxes, labels=None, cate=None, ques=None, dec_ids=None):
        eng = self._need_engine()
        dev = eng.device

        def dv(t, dtype):
            if t is None:
                return None
            return t.to(device=dev, dtype=dtype, non_blocking=True).contiguous()

        # torch raises IndexError for ids outside the embedding table (modeling_t5_our.py:196); on host tensors (what the
        # reference's collate_fn hands over) the check is free, device-resident batches are checked by the kernels (flag)
        V = self.config.vocab_size
        for nm, t, lo in (("input_ids", input_ids, 0), ("target_ids", labels, -100)):
            if t is not None and not t.is_cuda and t.numel():
                mn, mx = int(t.min()), int(t.max())
                if mx >= V or mn < lo or (lo == -100 and mn < 0 and bool(((t < 0) & (t != -100)).any())):
                    raise IndexError(f"{nm} holds ids outside [0, {V}) (min {mn}, max {mx})")
        ids = dv(input_ids, torch.int64)
        # fp32 RoI features (the reference's collate_fn) or bf16 ones from a packed shard (pipeline.py): same results, half the bytes
        feats = dv(vis_feats, torch.bfloat16 if vis_feats.dtype == torch.bfloat16 else torch.float32)
        bx = dv(boxes, torch.float32)
        lab = dv(labels, torch.int64)
        cate = dv(cate, torch.float32)
        ques = dv(ques, torch.float32)
        dec = dv(dec_ids, torch.int64)
        B, Lt = ids.shape
        N = feats.shape[1]
        assert feats.shape == (B, N, self.config.feat_dim), f"vis_feats {tuple(feats.shape)}"
        assert bx.shape == (B, N, 4), f"boxes {tuple(bx.shape)}"          # modeling_t5_our.py:105
        T = lab.shape[1] if lab is not None else (dec.shape[1] if dec is not None else 1)
        if dec is not None and lab is not None:
            assert dec.shape == lab.shape, "decoder_input_ids and labels must have the same shape"
        if cate is not None:
            assert cate.shape == (B, self.config.n_cate_classes) and ques.shape == (B, self.config.n_ques_classes)
        keep = (ids, feats, bx, lab, cate, ques, dec)
        cb = Engine.make_batch(B, Lt, N, T, feats, bx, ids, lab, cate, ques, dec)
        return cb, keep, (B, Lt, N, T)

    def _proto_state(self, proto_update, task_id=0, alpha=0.0, beta=0.0, memory=False):
        first = task_id not in self.Q_task_cur_proto
        has_mem = task_id in self.Q_task_mem_proto
        ps = CProtoState(Q_prototype=self._Q_prototype.data_ptr(), V_prototype=self._V_prototype.data_ptr(),
                         Q_num=self._Q_prototype_num.data_ptr(), V_num=self._V_prototype_num.data_ptr(),
                         proto_update=int(proto_update), task_id=int(task_id), first_step_of_task=int(first),
                         has_mem=int(has_mem), alpha=float(alpha or 0.0), beta=float(beta or 0.0),
                         memory_loss=int(bool(memory) and bool(proto_update)))
        if proto_update:
            # host-side image of the reference's dict bookkeeping (modeling_t5_our.py:467,476-485)
            if first:
                self.Q_task_cur_proto[task_id] = True
            elif task_id != 0:
                self.Q_task_mem_proto[task_id] = True
        return ps

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_world_size()
        return 1

    def _forward_native(self, cb, shape, proto_update, task_id, alpha, beta, training, memory=False):
        eng = self._engine
        B, Lt, N, T = shape
        eng.bind(B, Lt, N, T)
        self._step_seed += 1
        eng.forward_encoder(cb, self._base_seed * 2654435761 + self._step_seed, training)
        ps = self._proto_state(proto_update, task_id, alpha, beta, memory)
        sums_ready = False
        if proto_update and self.sync_prototypes and self._world() > 1:
            import torch.distributed as dist
            eng.proto_sums(cb)
            a = eng.L.vqacl_ws_offset(eng.h, b"curQ")
            z = eng.L.vqacl_ws_offset(eng.h, b"cntV") + self.config.n_cate_classes * 4
            dist.all_reduce(eng.ws[a:z].view(torch.float32), op=dist.ReduceOp.SUM)
            sums_ready = True
        # gradients are None (the reference sets them to None after every step, vqacl.py:486-487) => the coming backward starts
        # from zero and the arena can be cleared during the decoder forward
        prezero = training and self._grad_views[0][0].grad is None
        eng.forward_decoder(cb, ps, sums_ready, prezero_grads=prezero)

    def _native_backward(self, w_rows, gscale):
        eng = self._engine
        accumulate = self._grad_views[0][0].grad is not None
        if gscale is not None:
            gscale = gscale.detach().to(torch.float32).reshape(1).contiguous()     # device scalar, multiplied in ce_bwd
        world = self._world() if self.sync_grads else 1
        if world == 1:
            eng.backward(w_rows, accumulate, gscale=gscale)
        else:
            self._backward_overlapped(w_rows, accumulate, gscale)
        if not accumulate:
            for p, gv in self._grad_views:
                p.grad = gv

    def _backward_overlapped(self, w_rows, accumulate, gscale=None):
        """Gradient all-reduce (avg) over NCCL, bucketed by backward stage and issued on a side stream so it overlaps the
        remaining stages (replaces the reference's DDP reducer, which never fires — H11)."""
        import torch.distributed as dist
        eng = self._engine
        if accumulate:
            raise VqaclError("gradient accumulation across backward calls is not supported with multi-GPU sync")
        if self._comm_stream is None:
            # high priority: the collectives' CTAs must become resident as soon as an SM frees up, otherwise the back-to-back
            # persistent GEMMs (whose successor is already queued through PDL) starve them until backward is over
            self._comm_stream = torch.cuda.Stream(device=eng.device, priority=-1)
        main = torch.cuda.current_stream()
        n = eng.n_backward_stages()
        ranges = [eng.backward_stage_range(s) for s in range(n)]
        shard = self.shard_optimizer
        if shard:
            flush_after, buckets, tail = self.grad_shard_plan()
            world, rank = dist.get_world_size(), dist.get_rank()
        else:
            flush_after = plan_grad_buckets(ranges, self.grad_bucket_elems)
        n_sms = torch.cuda.get_device_properties(eng.device).multi_processor_count
        if self.comm_sms > 0:
            eng.set_gemm_sm_limit(n_sms - self.comm_sms)     # keep a few SMs free so the collectives' CTAs become resident

        def on_stage(s):
            # the engine has already made the communication stream wait for stage s on both of its streams
            with torch.cuda.stream(self._comm_stream):
                for a, b in flush_after.get(s, ()):
                    if shard:      # in place: this rank's slice of the bucket receives the average, the rest is scratch
                        k = (b - a) // world
                        dist.reduce_scatter_tensor(eng.G[a + rank * k:a + (rank + 1) * k], eng.G[a:b], op=dist.ReduceOp.AVG)
                    else:
                        dist.all_reduce(eng.G[a:b], op=dist.ReduceOp.AVG)
                if shard and s == n - 1:
                    dist.all_reduce(eng.G[tail[0]:tail[1]], op=dist.ReduceOp.AVG)

        eng.backward_overlapped(w_rows, False, self._comm_stream, on_stage, gscale=gscale)
        eng.set_gemm_sm_limit(0)
        main.wait_stream(self._comm_stream)

    def encode(self, input_ids, vis_inputs):
        """JointEncoder.forward (modeling_t5_our.py:175-339) in eval mode -> last_hidden_state [B, L+N, d] fp32."""
        cb, keep, (B, Lt, N, T) = self._stage_batch(input_ids, vis_inputs[0], vis_inputs[1])
        eng = self._engine
        eng.bind(B, Lt, N, T)
        eng.forward_encoder(cb, 0, False)
        return eng.ws_tensor("encoder_hidden_states", torch.float32, (B, Lt + N, self.config.d_model)).clone()

    def forward(self, input_ids=None, attention_mask=None, encoder_outputs=None, vis_inputs=None, vis_attention_mask=None,
                decoder_input_ids=None, decoder_attention_mask=None, past_key_values=None, use_cache=None, labels=None,
                inputs_embeds=None, decoder_inputs_embeds=None, head_mask=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, reduce_loss=False, return_hidden_state=False, **kwargs):
        """VLT5.forward (modeling_t5_our.py:514-713) for the teacher-forced call the reference makes (labels given).
        Returns per-row CE (`reduction='none'`, or the mean over valid rows with reduce_loss=True), bf16 logits (a view
        of engine memory: valid until the next backward / forward), encoder_hidden_states and the [B, L+N+2] memory mask."""
        if labels is None and decoder_input_ids is None:
            raise NotImplementedError("forward() needs labels (teacher forcing) or decoder_input_ids (logits of a given prefix); "
                                      "for decoding use generate() (KV-cached native loop)")
        for nm, v in (("encoder_outputs", encoder_outputs),
                      ("past_key_values", past_key_values), ("inputs_embeds", inputs_embeds),
                      ("decoder_inputs_embeds", decoder_inputs_embeds), ("head_mask", head_mask),
                      ("attention_mask", attention_mask), ("vis_attention_mask", vis_attention_mask),
                      ("decoder_attention_mask", decoder_attention_mask)):
            if v is not None:
                raise NotImplementedError(f"forward(): argument {nm} is not used by the VQACL train path and is not supported")
        proto_update = bool(kwargs.get("proto_update", False))
        # memory=True: the prototype pull losses. `memory_loss` is called at modeling_t5_our.py:591 but defined only in the
        # NExT-QA twin (nextqa/modeling_t5_nextqa.py:544-555); that definition is what runs here.
        memory = bool(kwargs.get("memory")) and proto_update
        cb, keep, shape = self._stage_batch(input_ids, vis_inputs[0], vis_inputs[1], labels,
                                            kwargs.get("cate_labels") if proto_update else None,
                                            kwargs.get("ques_labels") if proto_update else None, decoder_input_ids)
        self._forward_native(cb, shape, proto_update, kwargs.get("current_task_id", 0), kwargs.get("proto_alpha"),
                             kwargs.get("proto_beta"), self.training, memory)
        self._keep = keep
        out = self._collect_outputs(shape, keep[0])
        if labels is None:
            # decoder_input_ids without labels (modeling_t5_our.py:617-629, 659-671): logits of the given prefix, no loss. The
            # decoder re-runs on the whole prefix (no past_key_values); for greedy decoding generate() is the cached native loop.
            return out
        rows = self._engine.ws_tensor("loss_rows", torch.float32, (shape[0] * shape[3],))
        if memory:
            mem2 = self._engine.ws_tensor("loss_memory", torch.float32, (2,))
            loss, out.loss_memory_Q, out.loss_memory_V = _StepLossMem.apply(self._anchor, self, rows, None, mem2)
        else:
            loss = _RowLoss.apply(self._anchor, self, rows)
        if reduce_loss:
            loss = loss.sum() / (keep[3].view(-1) != -100).sum().clamp(min=1)
        out.loss = loss
        return out

    def _collect_outputs(self, shape, ids):
        eng, c = self._engine, self.config
        B, Lt, N, T = shape
        S = Lt + N
        ldv = (c.vocab_size + 255) // 256 * 256
        out = VLSeq2SeqLMOutput()
        out.logits = eng.ws_tensor("logits", torch.bfloat16, (B, T, c.vocab_size), pitch=ldv)
        out.encoder_hidden_states = eng.ws_tensor("encoder_hidden_states", torch.float32, (B, S, c.d_model))
        out.encoder_attention_mask = eng.ws_tensor("encoder_attention_mask", torch.float32, (B, S + 2))
        out.max_idx_Q = eng.ws_tensor("idxQ", torch.int64, (B,))
        out.max_idx_V = eng.ws_tensor("idxV", torch.int64, (B,))
        return out

    @torch.no_grad()
    def generate(self, input_ids=None, vis_inputs=None, max_length=20, num_beams=1, **kwargs):
        """Greedy search exactly as the reference reaches it (num_beams is parsed but never forwarded).
        Returns int64 [B, <= max_length] with the start token in column 0."""
        if num_beams not in (None, 1):
            raise NotImplementedError("only greedy search is on the reference's path (vqa_model.py:112-116)")
        cb, keep, (B, Lt, N, T) = self._stage_batch(input_ids, vis_inputs[0], vis_inputs[1])
        eng = self._engine
        eng.bind(B, Lt, N, 1)
        ps = self._proto_state(False)
        return eng.generate(cb, ps, int(max_length))


class VLT5VQA(VLT5):
    """VL-T5/src/vqa_model.py:10-121."""

    def __init__(self, config, num_answers=None, label2ans=None):
        super().__init__(config)
        self.num_answers = num_answers
        self.label2ans = label2ans

    def train_step(self, batch, current_task_id, proto_alpha, proto_beta, mem_num_Q=0, total_num_Q=1000, memory=False):
        """vqa_model.py:18-65: H2D of the batch, forward with proto_update=True, per-sample masked mean x soft score,
        batch mean. `mem_num_Q`, `total_num_Q` are accepted and ignored, as in the reference."""
        eng = self._need_engine()
        cb, keep, shape = self._stage_batch(batch["input_ids"], batch["vis_feats"], batch["boxes"], batch["target_ids"],
                                            batch["cate_labels"], batch["ques_labels"])
        B, Lt, N, T = shape
        scores = batch["scores"].to(device=eng.device, dtype=torch.float32, non_blocking=True).contiguous()
        self._forward_native(cb, shape, True, current_task_id, proto_alpha, proto_beta, self.training, memory)
        w_rows = torch.empty(B * T, dtype=torch.float32, device=eng.device)
        eng.loss_tail(keep[3], scores, B, T, self._loss_buf, w_rows)
        self._keep = keep + (scores,)
        out = self._collect_outputs(shape, keep[0])
        result = {"encoder_hidden_states": out.encoder_hidden_states, "BL": (B, T),
                  "encoder_attention_mask": out.encoder_attention_mask, "logits": out.logits,
                  "max_idx_Q": out.max_idx_Q, "max_idx_V": out.max_idx_V}
        if memory:
            # The reference's train_step looks for a key 'loss_memory' that its forward never sets (vqa_model.py:61-62; the
            # comment there names the intended value); the tuple is provided so that Trainer.train_step's
            # `loss + lambda_Q * loss_memory_Q + lambda_V * loss_memory_V` (vqacl.py:448-450) takes effect.
            mem2 = eng.ws_tensor("loss_memory", torch.float32, (2,))
            result["loss"], lq, lv = _StepLossMem.apply(self._anchor, self, self._loss_buf[:1], w_rows, mem2)
            result["loss_memory"] = (lq, lv)
        else:
            result["loss"] = _StepLoss.apply(self._anchor, self, self._loss_buf[:1], w_rows)
        return result

    @torch.no_grad()
    def test_step(self, batch, **kwargs):
        """vqa_model.py:68-121 (generation branch; the classifier branch is dead in the reference)."""
        self.eval()
        output = self.generate(input_ids=batch["input_ids"], vis_inputs=(batch["vis_feats"], batch["boxes"]), **kwargs)
        result = {"token_ids": output}
        tok = getattr(self, "tokenizer", None)
        if tok is not None:
            result["pred_ans"] = tok.batch_decode(output, skip_special_tokens=True)
        else:
            result["pred_ans"] = [" ".join(str(int(t)) for t in row if int(t) > 1) for row in output.cpu()]
        return result
