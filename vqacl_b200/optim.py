"""Step tail of Trainer.train_step (VL-T5/src/vqacl.py:466-487; optimizer factory trainer_base.py:130-198):
`clip_grad_norm_(params, 5.0)` + transformers-4.2.1 `AdamW(lr, eps=adam_eps, correct_bias=True)` +
`get_constant_schedule_with_warmup`, as ONE multi-tensor kernel pair over the engine's flat arena
(grad sum-of-squares, then clip * Adam * decoupled decay * bf16 refresh; csrc/optim.cu).

`FusedAdamW` is a torch.optim.Optimizer, so the reference's LambdaLR scheduler and its per-category-group
re-creation (vqacl.py:324-329) work unchanged.
"""
import torch

from ._lib import VqaclError
from .modeling import chunk_events, owned_slices


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, max_grad_norm=0.0,
                 overlap_with_next_forward=False, shard_state=None):
        """overlap_with_next_forward: run the (HBM-bound) update on a side stream, in the order the next forward reads the
        parameters; `train_step` / `forward` / `generate` / `state_dict` wait for exactly what they need. Code that reads
        parameter tensors directly right after `step()` must call `model.param_sync()` first (off by default).
        shard_state (None = automatically with more than one rank and model.sync_grads): partition the optimizer over the
        ranks. Backward then reduce-scatters the gradients of the GEMM matrices (half the wire bytes of an all-reduce), each
        rank keeps Adam moments for, and updates, only its 1/N of them, and the refreshed bf16 weights are all-gathered on the
        communication stream while the next forward already runs (it waits chunk by chunk). The small fp32-read tail of the
        arena (embedding table, norm weights, biases) is all-reduced and updated everywhere. `p.grad` / `p.data` of the GEMM
        matrices are then only partially current on each rank: `model.state_dict()` / `model.gather_params()` all-gather the
        fp32 masters."""
        model = getattr(model, "module", model)
        eng = model._need_engine()
        no_decay = ["bias", "LayerNorm.weight"]                     # trainer_base.py:148 (T5 'layer_norm.weight' does NOT match)
        named = list(model.named_parameters())
        groups = [
            {"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": weight_decay},
            {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
        ]
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        # the arena's decay boundary must agree with the name rule above
        for n, p in named:
            key = n
            off, rows, cols, grp = eng.table[key]
            decays = not any(nd in n for nd in no_decay)
            if grp != 2 and decays != (grp == 0):
                raise VqaclError(f"arena decay group of {n} disagrees with the reference's no_decay rule")
        self.model, self.eng = model, eng
        self.max_grad_norm = float(max_grad_norm)
        self.overlap = bool(overlap_with_next_forward)
        world = model._world()
        self.shard = bool(world > 1 and model.sync_grads) if shard_state is None else bool(shard_state)
        if self.shard and world == 1:
            self.shard = False
        if self.shard:
            import torch.distributed as dist
            try:
                _, buckets, tail = model.grad_shard_plan()
            except VqaclError:
                self.shard = False          # world size that does not divide the buckets into aligned slices: replicate
        if not self.shard:
            model.gather_params()       # a previous sharded optimizer left only this rank's fp32 master slices current
        model.shard_optimizer = self.shard
        if self.shard:
            self.owned = owned_slices(buckets, world, dist.get_rank())
            self.buckets, self.tail = buckets, tail
            n_state = sum(b - a for a, b in self.owned) + (tail[1] - tail[0])
        else:
            n_state = eng.n_train
        self.exp_avg = torch.zeros(n_state, dtype=torch.float32, device=eng.device)
        self.exp_avg_sq = torch.zeros(n_state, dtype=torch.float32, device=eng.device)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float32, device=eng.device)
        self.t = 0

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        """One optimizer step over every trained parameter. `max_grad_norm` > 0 fuses clip_grad_norm_ (vqacl.py:475-476)
        into the same pass; after the call `self.grad_sumsq` holds the squared global gradient norm (device, no sync)."""
        g0, g1 = self.param_groups
        if g0["lr"] != g1["lr"] or g1["weight_decay"] != 0.0:
            raise VqaclError("FusedAdamW: both groups share one lr and only group 0 decays (trainer_base.py:148-160)")
        if self.model._grad_views[0][0].grad is None:
            return None        # HF AdamW skips parameters without gradients: nothing to do before the first backward
        self.t += 1
        b1, b2 = g0["betas"]
        mg = self.max_grad_norm if max_grad_norm is None else float(max_grad_norm)
        if self.shard:
            return self._step_sharded(float(g0["lr"]), float(b1), float(b2), float(g0["eps"]), float(g0["weight_decay"]), mg)
        self.eng.clip_adamw(self.exp_avg, self.exp_avg_sq, float(g0["lr"]), float(b1), float(b2), float(g0["eps"]),
                            float(g0["weight_decay"]), self.t, mg, self.grad_sumsq, overlap=self.overlap)
        return None

    def _step_sharded(self, lr, b1, b2, eps, wd, mg):
        import torch.distributed as dist
        eng, model = self.eng, self.model
        rank = dist.get_rank()
        # global gradient norm: every rank sums its own slices (rank 0 adds the replicated tail once), then one scalar all-reduce
        eng.grad_sumsq_ranges(self.owned + ([self.tail] if rank == 0 else []), self.grad_sumsq)
        dist.all_reduce(self.grad_sumsq, op=dist.ReduceOp.SUM)
        off = 0
        for a, b in self.owned:
            eng.adamw_range(self.exp_avg[off:], self.exp_avg_sq[off:], a, b, lr, b1, b2, eps, wd, self.t, self.grad_sumsq, mg)
            off += b - a
        eng.adamw_range(self.exp_avg[off:], self.exp_avg_sq[off:], self.tail[0], self.tail[1], lr, b1, b2, eps, wd, self.t,
                        self.grad_sumsq, mg)
        # all-gather the refreshed bf16 weights in the order the next forward reads them (the reverse of backward's bucket
        # order) on the communication stream; the forward waits per parameter chunk (vqacl_set_param_events)
        if model._comm_stream is None:
            model._comm_stream = torch.cuda.Stream(device=eng.device, priority=-1)
        comm = model._comm_stream
        comm.wait_stream(torch.cuda.current_stream())
        order = list(range(len(self.buckets) - 1, -1, -1))
        evs = []
        with torch.cuda.stream(comm):
            for i in order:
                (a, b), (oa, ob) = self.buckets[i], self.owned[i]
                dist.all_gather_into_tensor(eng.W[a:b], eng.W[oa:ob])
                ev = torch.cuda.Event()
                ev.record(comm)
                evs.append(ev)
        need = chunk_events(eng.param_chunks(), [self.buckets[i] for i in order])
        model._param_events = [evs[k] if k is not None else None for k in need]
        eng.set_param_events(model._param_events)
        model._masters_stale = True
        if not self.overlap:
            torch.cuda.current_stream().wait_stream(comm)
        return None

    def zero_grad(self, set_to_none=True):
        """vqacl.py:486-487 sets grads to None; the next backward re-attaches the arena views."""
        for p, _ in self.model._grad_views:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()


def get_constant_schedule_with_warmup(optimizer, num_warmup_steps, last_epoch=-1):
    """transformers.get_constant_schedule_with_warmup (trainer_base.py:189-190)."""

    def lr_lambda(step):
        if step < num_warmup_steps:
            return float(step) / float(max(1.0, num_warmup_steps))
        return 1.0

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda, last_epoch)
