"""Step tail of Trainer.train_step (VL-T5/src/vqacl.py:466-487; optimizer factory trainer_base.py:130-198):
`clip_grad_norm_(params, 5.0)` + transformers-4.2.1 `AdamW(lr, eps=adam_eps, correct_bias=True)` +
`get_constant_schedule_with_warmup`, as ONE multi-tensor kernel pair over the engine's flat arena
(grad sum-of-squares, then clip * Adam * decoupled decay * bf16 refresh; csrc/optim.cu).

`FusedAdamW` is a torch.optim.Optimizer, so the reference's LambdaLR scheduler and its per-category-group
re-creation (vqacl.py:324-329) work unchanged.
"""
import torch

from ._lib import VqaclError


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, max_grad_norm=0.0,
                 overlap_with_next_forward=False):
        """overlap_with_next_forward: run the (HBM-bound) update on a side stream, in the order the next forward reads the
        parameters; `train_step` / `forward` / `generate` / `state_dict` wait for exactly what they need. Code that reads
        parameter tensors directly right after `step()` must call `model.param_sync()` first (off by default)."""
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
        self.exp_avg = torch.zeros(eng.n_train, dtype=torch.float32, device=eng.device)
        self.exp_avg_sq = torch.zeros(eng.n_train, dtype=torch.float32, device=eng.device)
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
        self.eng.clip_adamw(self.exp_avg, self.exp_avg_sq, float(g0["lr"]), float(b1), float(b2), float(g0["eps"]),
                            float(g0["weight_decay"]), self.t, mg, self.grad_sumsq, overlap=self.overlap)
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
