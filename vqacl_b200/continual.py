"""Host-side continual-learning state around the hot path: the rehearsal memory of
Trainer.train (VL-T5/src/vqacl.py:169-203) and a checkpoint that carries everything a run needs to RESUME — the reference
saves only the weights per task (`<task>_LAST.pth`, trainer_base.py:246-249) and the two prototype banks once at the very end
(vqacl.py:419-423), so a job that dies mid-sequence restarts without its SI banks, their per-task bookkeeping or its
exemplar memory. Rehearsal sampling and the task scheduler stay in Python, as BASELINE.json's north_star asks.
"""
import random

import torch


class RehearsalMemory:
    """`Examplar_set` of the reference: per object-category group, one list of exemplars per finished task."""

    def __init__(self, category_splits, M=5000):
        self.splits = {g: list(v) for g, v in category_splits.items()}
        self.M = int(M)
        self.examplar_set = {g: [] for g in self.splits}

    def grow(self, task_idx, prev_task_items, img_cate_map, rng=random):
        """Entering task `task_idx` (>= 1): add exemplars of task `task_idx - 1` and shrink the older tasks' shares so the
        whole memory stays within M. `prev_task_items`: the previous task's training records (dicts with 'img_id'); shuffled
        in place with `rng` exactly like the reference. Returns (All_examplar, each_memory). vqacl.py:171-203."""
        each_memory = int(self.M / task_idx)                                   # :171
        rng.shuffle(prev_task_items)                                           # :176
        each_memory_for_cate = int(each_memory / len(self.splits))             # :177
        for cate in self.splits:                                               # :179-189
            num = 0
            self.examplar_set[cate].append([])
            for d in prev_task_items:
                img_id = d["img_id"]
                if img_id in img_cate_map and img_cate_map[img_id] in self.splits[cate]:
                    self.examplar_set[cate][task_idx - 1].append(d)
                    num += 1
                    if num >= each_memory_for_cate:
                        break
        for cate in self.splits:                                               # :193-195
            for i in range(task_idx):
                self.examplar_set[cate][i] = self.examplar_set[cate][i][:each_memory_for_cate]
        return self.all_examplars(), each_memory

    def all_examplars(self):
        out = []
        for g in self.examplar_set:                                            # :197-200
            for task_set in self.examplar_set[g]:
                out += task_set
        return out

    def state_dict(self):
        return {"M": self.M, "splits": self.splits, "examplar_set": self.examplar_set}

    def load_state_dict(self, sd):
        self.M, self.splits, self.examplar_set = sd["M"], sd["splits"], sd["examplar_set"]


def save_training_state(path, model, task_idx, memory=None, extra=None):
    """Everything needed to continue a VQACL run after `task_idx` finished: weights under the DDP-style 'module.' keys the
    reference writes (trainer_base.py:246-249), both SI banks with their running counts and the per-task bookkeeping of
    update_prototype (which tasks have a current / memory prototype, modeling_t5_our.py:467-485), the rehearsal memory and the
    host RNG states that drive group shuffling (vqacl.py:314) and exemplar selection."""
    model = getattr(model, "module", model)
    state = {
        "model": {"module." + k: v.detach().cpu() for k, v in model.state_dict().items()},
        "task_idx": int(task_idx),
        "Q_prototype": model.Q_prototype.detach().cpu(), "V_prototype": model.V_prototype.detach().cpu(),
        "Q_prototype_num": model.Q_prototype_num.detach().cpu(), "V_prototype_num": model.V_prototype_num.detach().cpu(),
        "Q_task_cur_proto": sorted(model.Q_task_cur_proto), "Q_task_mem_proto": sorted(model.Q_task_mem_proto),
        "step_seed": int(model._step_seed),
        "memory": memory.state_dict() if memory is not None else None,
        "rng": {"python": random.getstate(), "torch": torch.get_rng_state()},
        "extra": extra,
    }
    torch.save(state, path)
    return path


def load_training_state(path, model, memory=None, restore_rng=True):
    """Inverse of save_training_state; returns (task_idx, extra). The model may be packed on its CUDA device already."""
    model = getattr(model, "module", model)
    state = torch.load(path, weights_only=False)
    model.load_state_dict(state["model"], strict=True)
    model.Q_prototype, model.V_prototype = state["Q_prototype"], state["V_prototype"]
    if model.Q_prototype_num is not None:
        model.Q_prototype_num.copy_(state["Q_prototype_num"])
        model.V_prototype_num.copy_(state["V_prototype_num"])
    else:
        model._Q_prototype_num, model._V_prototype_num = state["Q_prototype_num"].clone(), state["V_prototype_num"].clone()
    model.Q_task_cur_proto = {t: True for t in state["Q_task_cur_proto"]}
    model.Q_task_mem_proto = {t: True for t in state["Q_task_mem_proto"]}
    model._step_seed = state["step_seed"]
    if memory is not None and state["memory"] is not None:
        memory.load_state_dict(state["memory"])
    if restore_rng:
        random.setstate(state["rng"]["python"])
        torch.set_rng_state(state["rng"]["torch"])
    return state["task_idx"], state["extra"]
