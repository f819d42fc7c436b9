"""ctypes host of the native step engine (include/vqacl_b200.h).

PyTorch is plumbing only: it owns the device memory (flat parameter / gradient / bf16 arenas, the activation
workspace), the CUDA stream and — for N > 1 — the NCCL process group. Every kernel of the step is launched from C++
by libvqacl_b200.so; there is no eager/PyTorch fallback (a missing library or a failing call raises VqaclError).
"""
import ctypes
import math
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int64, c_uint32, c_void_p

import torch

from ._lib import VqaclError, check, cur_stream, lib, ptr


class CConfig(Structure):
    """vqacl_config (include/vqacl_b200.h)."""
    _fields_ = [
        ("vocab_size", c_int), ("d_model", c_int), ("d_kv", c_int), ("n_heads", c_int), ("d_ff", c_int),
        ("n_enc_layers", c_int), ("n_dec_layers", c_int), ("n_buckets", c_int), ("feat_dim", c_int),
        ("n_images", c_int), ("n_ques", c_int), ("n_cate", c_int), ("split_L", c_int),
        ("pad_id", c_int), ("eos_id", c_int), ("start_id", c_int), ("eps", c_float), ("dropout", c_float),
    ]


class CBatch(Structure):
    """vqacl_batch."""
    _fields_ = [
        ("B", c_int), ("L", c_int), ("N", c_int), ("T", c_int),
        ("vis_feats", c_void_p), ("boxes", c_void_p), ("input_ids", c_void_p), ("labels", c_void_p),
        ("cate_labels", c_void_p), ("ques_labels", c_void_p), ("decoder_input_ids", c_void_p), ("vis_feats_bf16", c_void_p),
    ]


class CProtoState(Structure):
    """vqacl_proto_state."""
    _fields_ = [
        ("Q_prototype", c_void_p), ("V_prototype", c_void_p), ("Q_num", c_void_p), ("V_num", c_void_p),
        ("proto_update", c_int), ("task_id", c_int), ("first_step_of_task", c_int), ("has_mem", c_int),
        ("alpha", c_float), ("beta", c_float), ("memory_loss", c_int),
    ]


_SIGS_DONE = False
STAGE_CB = ctypes.CFUNCTYPE(None, c_int, c_void_p)


def _declare(L):
    global _SIGS_DONE
    if _SIGS_DONE:
        return
    L.vqacl_engine_create.argtypes = [POINTER(CConfig), POINTER(c_void_p)]
    L.vqacl_engine_destroy.argtypes = [c_void_p]
    L.vqacl_engine_destroy.restype = None
    L.vqacl_param_count.argtypes = [c_void_p]
    L.vqacl_param_info.argtypes = [c_void_p, c_int, c_char_p, c_int, POINTER(c_int64), POINTER(c_int), POINTER(c_int),
                                   POINTER(c_int)]
    L.vqacl_arena_elems.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64)]
    L.vqacl_arena_elems.restype = c_int64
    L.vqacl_bind_arena.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    L.vqacl_set_rel_buckets.argtypes = [c_void_p, c_void_p, c_void_p]
    L.vqacl_refresh_bf16.argtypes = [c_void_p, c_void_p]
    L.vqacl_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int, c_int]
    L.vqacl_workspace_bytes.restype = c_int64
    L.vqacl_bind_workspace.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int]
    L.vqacl_ws_offset.argtypes = [c_void_p, c_char_p]
    L.vqacl_ws_offset.restype = c_int64
    L.vqacl_forward_encoder.argtypes = [c_void_p, POINTER(CBatch), POINTER(CProtoState), c_uint32, c_int, c_void_p]
    L.vqacl_forward_decoder.argtypes = [c_void_p, POINTER(CBatch), POINTER(CProtoState), c_int, c_int, c_void_p]
    L.vqacl_proto_sums.argtypes = [c_void_p, POINTER(CBatch), c_void_p]
    L.vqacl_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    L.vqacl_backward_overlapped.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, STAGE_CB, c_void_p, c_void_p]
    L.vqacl_set_memory_loss_grads.argtypes = [c_void_p, c_void_p]
    L.vqacl_backward_stages.argtypes = [c_void_p]
    L.vqacl_backward_stage_range.argtypes = [c_void_p, c_int, POINTER(c_int64), POINTER(c_int64)]
    L.vqacl_loss_tail.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.vqacl_clip_adamw.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float, c_int,
                                   c_float, c_void_p, c_int, c_void_p]
    L.vqacl_param_sync.argtypes = [c_void_p, c_void_p]
    L.vqacl_arena_tail.argtypes = [c_void_p]
    L.vqacl_arena_tail.restype = c_int64
    L.vqacl_set_param_events.argtypes = [c_void_p, POINTER(c_void_p), c_int]
    L.vqacl_collate_device.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                       c_void_p]
    L.vqacl_device_errors.argtypes = [c_void_p, POINTER(c_int), c_void_p]
    L.vqacl_grad_sumsq_ranges.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), c_int, c_void_p, c_void_p]
    L.vqacl_adamw_range.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_float, c_float, c_float,
                                    c_int, c_void_p, c_float, c_void_p]
    L.vqacl_generate.argtypes = [c_void_p, POINTER(CBatch), POINTER(CProtoState), c_int, c_void_p, c_void_p, c_int64,
                                 POINTER(c_int), c_void_p]
    L.vqacl_generate_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int, c_int]
    L.vqacl_generate_workspace_bytes.restype = c_int64
    L.vqacl_set_gemm_sm_limit.argtypes = [c_int]
    L.vqacl_launch_count.argtypes = []
    L.vqacl_launch_count.restype = c_int64
    _SIGS_DONE = True


def relative_position_bucket_host(rel, bidirectional, num_buckets=32, max_distance=128):
    """Bucket of one relative position (key - query), the T5 formula (HF T5Attention._relative_position_bucket,
    hf5.5 modeling_t5.py:189-234; identical in 4.2.1). Evaluated through torch in fp32 so that the float->int
    truncation of the log branch matches the reference's tensor arithmetic bit for bit."""
    rp = torch.tensor([rel], dtype=torch.long)
    out = torch.zeros_like(rp)
    nb = num_buckets
    if bidirectional:
        nb //= 2
        out = out + (rp > 0).long() * nb
        rp = rp.abs()
    else:
        rp = -torch.min(rp, torch.zeros_like(rp))
    max_exact = nb // 2
    small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return int((out + torch.where(small, rp, large)).item())


def rel_bucket_table(bidirectional, num_buckets=32, max_distance=128, span=64):
    """int32[2*span-1]: bucket of (key - query + span - 1) for key, query in [0, span)."""
    return torch.tensor([relative_position_bucket_host(r, bidirectional, num_buckets, max_distance)
                         for r in range(-(span - 1), span)], dtype=torch.int32)


class Engine:
    """One native engine per model replica (one process per GPU)."""

    def __init__(self, cfg, device):
        L = lib()
        _declare(L)
        self.L = L
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise VqaclError("vqacl_b200 has no CPU path: the engine needs a CUDA device (sm_100a)")
        self.cfg = cfg
        self.c = CConfig(
            vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, n_heads=cfg.num_heads, d_ff=cfg.d_ff,
            n_enc_layers=cfg.num_layers, n_dec_layers=cfg.num_decoder_layers, n_buckets=cfg.relative_attention_num_buckets,
            feat_dim=cfg.feat_dim, n_images=cfg.n_images, n_ques=cfg.n_ques_classes, n_cate=cfg.n_cate_classes,
            split_L=cfg.proto_split_L, pad_id=cfg.pad_token_id, eos_id=cfg.eos_token_id,
            start_id=cfg.decoder_start_token_id, eps=cfg.layer_norm_epsilon, dropout=cfg.dropout_rate)
        h = c_void_p()
        check(L.vqacl_engine_create(byref(self.c), byref(h)))
        self.h = h
        # parameter table
        self.table = {}
        buf = ctypes.create_string_buffer(256)
        for i in range(L.vqacl_param_count(h)):
            off, rows, cols, grp = c_int64(), c_int(), c_int(), c_int()
            check(L.vqacl_param_info(h, i, buf, 256, byref(off), byref(rows), byref(cols), byref(grp)))
            self.table[buf.value.decode()] = (off.value, rows.value, cols.value, grp.value)
        nd, nt = c_int64(), c_int64()
        self.n_total = L.vqacl_arena_elems(h, byref(nd), byref(nt))
        self.n_decay, self.n_train = nd.value, nt.value
        self.n_tail = L.vqacl_arena_tail(h)      # [0, n_tail): GEMM-only matrices; [n_tail, n_train): fp32-read parameters
        with torch.cuda.device(self.device):
            self.P = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
            self.G = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
            self.W = torch.zeros(self.n_total, dtype=torch.bfloat16, device=self.device)
            check(L.vqacl_bind_arena(h, ptr(self.P), ptr(self.G), ptr(self.W)))
            nb = cfg.relative_attention_num_buckets
            md = getattr(cfg, "relative_attention_max_distance", 128)
            self.enc_bucket = rel_bucket_table(True, nb, md).contiguous()      # host tables (kernel parameters)
            self.dec_bucket = rel_bucket_table(False, nb, md).contiguous()
            check(L.vqacl_set_rel_buckets(h, ptr(self.enc_bucket), ptr(self.dec_bucket)))
        self.ws = None
        self.ws_shape = None
        self.gen_ws = None
        self.bf16_stale = True

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.vqacl_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- views ---------------------------------------------------------------------------------------------------
    def param_view(self, name, arena=None):
        off, rows, cols, _ = self.table[name]
        a = self.P if arena is None else arena
        v = a[off:off + rows * cols]
        return v.view(rows, cols)

    def refresh_bf16(self):
        check(self.L.vqacl_refresh_bf16(self.h, cur_stream()))
        self.bf16_stale = False

    # ---- workspace -----------------------------------------------------------------------------------------------
    def bind(self, B, Lt, N, T):
        shape = (B, Lt, N, T)
        if shape == self.ws_shape:
            return
        need = self.L.vqacl_workspace_bytes(self.h, B, Lt, N, T)
        if need < 0:
            check(1)
        if self.ws is None or self.ws.numel() < need:
            self.ws = None
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        check(self.L.vqacl_bind_workspace(self.h, ptr(self.ws), self.ws.numel(), B, Lt, N, T))
        self.ws_shape = shape

    def ws_tensor(self, name, dtype, shape, pitch=None):
        """Typed view of a named workspace region (outputs the reference returns to its caller)."""
        off = self.L.vqacl_ws_offset(self.h, name.encode())
        if off < 0:
            raise VqaclError(f"unknown workspace region {name}")
        es = torch.empty((), dtype=dtype).element_size()
        rows = int(math.prod(shape[:-1])) if len(shape) > 1 else 1
        cols = shape[-1]
        p = cols if pitch is None else pitch
        flat = self.ws[off:off + rows * p * es].view(dtype)
        return flat.view(rows, p)[:, :cols].reshape(*shape) if pitch is not None else flat.view(*shape)

    # ---- step ----------------------------------------------------------------------------------------------------
    @staticmethod
    def make_batch(B, Lt, N, T, feats, boxes, ids, labels=None, cate=None, ques=None, dec_ids=None):
        bf = feats.dtype == torch.bfloat16      # packed feature shards hand the GEMM operand format over directly
        return CBatch(B=B, L=Lt, N=N, T=T, vis_feats=None if bf else feats.data_ptr(), vis_feats_bf16=feats.data_ptr() if bf else None,
                      boxes=boxes.data_ptr(), input_ids=ids.data_ptr(),
                      labels=labels.data_ptr() if labels is not None else None,
                      cate_labels=cate.data_ptr() if cate is not None else None,
                      ques_labels=ques.data_ptr() if ques is not None else None,
                      decoder_input_ids=dec_ids.data_ptr() if dec_ids is not None else None)

    def forward_encoder(self, cb, seed, training):
        if self.bf16_stale:
            self.refresh_bf16()
        check(self.L.vqacl_forward_encoder(self.h, byref(cb), None, seed & 0xFFFFFFFF, int(training), cur_stream()))

    def proto_sums(self, cb):
        check(self.L.vqacl_proto_sums(self.h, byref(cb), cur_stream()))

    def forward_decoder(self, cb, ps, sums_ready=False, prezero_grads=False):
        check(self.L.vqacl_forward_decoder(self.h, byref(cb), byref(ps), int(sums_ready), int(prezero_grads), cur_stream()))

    def loss_tail(self, labels, scores, B, T, loss_out, w_rows):
        lr = self.ws_tensor("loss_rows", torch.float32, (B * T,))
        check(self.L.vqacl_loss_tail(ptr(lr), ptr(labels), ptr(scores), B, T, ptr(loss_out), ptr(w_rows), cur_stream()))

    def set_param_events(self, events):
        """events: one torch.cuda.Event (or None) per parameter chunk, or [] to clear."""
        n = len(events)
        arr = (c_void_p * max(n, 1))(*[c_void_p(ev.cuda_event) if ev is not None else c_void_p(0) for ev in events])
        check(self.L.vqacl_set_param_events(self.h, arr, n))

    def param_chunks(self):
        """Arena ranges of the parameter chunks the forward waits for, in forward order (engine.cu: vqacl_clip_adamw)."""
        t = self.table
        Le, Ld = self.cfg.num_layers, self.cfg.num_decoder_layers
        oWf = t["encoder.visual_embedding.feat_embedding.0.weight"][0]
        enc = [t[f"encoder.block.{l}.layer.0.SelfAttention.q.weight"][0] for l in range(Le)]
        out = [(oWf, self.n_train)]
        for l in range(Le):
            out.append((enc[l], enc[l + 1] if l + 1 < Le else oWf))
        out.append((0, enc[0] if Le else oWf))
        return out

    def set_memory_loss_grads(self, g2):
        """g2: device fp32[2] = d loss / d (loss_memory_Q, loss_memory_V) for the next backward, or None."""
        check(self.L.vqacl_set_memory_loss_grads(self.h, ptr(g2)))

    def n_backward_stages(self):
        return self.L.vqacl_backward_stages(self.h)

    def backward_stage_range(self, stage):
        a, b = c_int64(), c_int64()
        check(self.L.vqacl_backward_stage_range(self.h, stage, byref(a), byref(b)))
        return a.value, b.value

    def backward(self, w_rows, accumulate=False, stage_begin=0, stage_end=-1, gscale=None):
        check(self.L.vqacl_backward(self.h, ptr(w_rows), ptr(gscale), int(accumulate), stage_begin, stage_end, cur_stream()))

    def backward_overlapped(self, w_rows, accumulate, comm_stream, on_stage, gscale=None):
        """One-call backward; `on_stage(stage)` runs on the host right after the engine ordered `comm_stream` behind that stage."""
        errors = []

        def _cb(stage, _user):
            try:
                on_stage(stage)
            except BaseException as e:     # exceptions cannot cross the C frame: re-raised below
                errors.append(e)
        cb = STAGE_CB(_cb)
        rc = self.L.vqacl_backward_overlapped(self.h, ptr(w_rows), ptr(gscale), int(accumulate), c_void_p(comm_stream.cuda_stream), cb, None,
                                              cur_stream())
        if errors:
            raise errors[0]
        check(rc)

    def clip_adamw(self, m, v, lr, beta1, beta2, eps, wd, step, max_norm, norm_out=None, overlap=False):
        """norm_out (fp32[1], device) receives the SQUARED global gradient norm."""
        check(self.L.vqacl_clip_adamw(self.h, ptr(m), ptr(v), lr, beta1, beta2, eps, wd, step, max_norm,
                                      ptr(norm_out), int(overlap), cur_stream()))

    def param_sync(self):
        """Order a pending overlapped optimizer step before the current stream (needed before touching parameters outside
        the engine's own forward / generate calls)."""
        check(self.L.vqacl_param_sync(self.h, cur_stream()))
        # bf16_stale is NOT touched here: the fused optimizer refreshes the bf16 copies itself and never marks them stale;
        # the flag only records fp32 edits made outside the engine (load_state_dict / apply / _pack), which need a refresh

    def check_device_errors(self):
        """Raise for errors that kernels could only flag (synchronises the current stream)."""
        f = c_int(0)
        check(self.L.vqacl_device_errors(self.h, byref(f), cur_stream()))
        if f.value & 1:
            raise IndexError("token id outside [0, vocab_size) reached an embedding gather (input_ids / target_ids); "
                             "was resize_token_embeddings() skipped?")

    def grad_sumsq_ranges(self, ranges, out):
        n = len(ranges)
        b = (c_int64 * n)(*[r[0] for r in ranges])
        e = (c_int64 * n)(*[r[1] for r in ranges])
        check(self.L.vqacl_grad_sumsq_ranges(self.h, b, e, n, ptr(out), cur_stream()))

    def adamw_range(self, m, v, begin, end, lr, beta1, beta2, eps, wd, step, sumsq, max_norm):
        check(self.L.vqacl_adamw_range(self.h, ptr(m), ptr(v), begin, end, lr, beta1, beta2, eps, wd, step, ptr(sumsq), max_norm,
                                       cur_stream()))

    def generate(self, cb, ps, max_len):
        B = cb.B
        need = self.L.vqacl_generate_workspace_bytes(self.h, B, cb.L, cb.N, max_len)
        if self.gen_ws is None or self.gen_ws.numel() < need:
            self.gen_ws = None
            self.gen_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if self.bf16_stale:
            self.refresh_bf16()
        out = torch.zeros(B, max_len, dtype=torch.int64, device=self.device)
        n = c_int(0)
        check(self.L.vqacl_generate(self.h, byref(cb), byref(ps), max_len, ptr(out), ptr(self.gen_ws), self.gen_ws.numel(),
                                    byref(n), cur_stream()))
        self.check_device_errors()       # generate has synchronised anyway
        return out[:, :n.value]

    def set_gemm_sm_limit(self, n):
        check(self.L.vqacl_set_gemm_sm_limit(int(n)))

    def launch_count(self):
        return self.L.vqacl_launch_count()
