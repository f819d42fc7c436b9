"""The VQACL outer loop (VL-T5/src/vqacl.py:146-426) driving the B200 path on synthetic data.

Shape of the reference's training procedure, kept in Python as BASELINE.json's north_star asks (task scheduler and
rehearsal-memory sampling are host logic): for each of the question-type tasks, for each object-category group in a
shuffled order, a NEW optimizer + warm-up schedule is created (vqacl.py:324-329) and `epochs` passes are made over the
group's loader; from the second task on every new-task step is followed by a step on a rehearsal batch drawn from a
memory of earlier tasks (`zip(train_loader, cycle(memory_loader))`, vqacl.py:358-373) with the SAME current_task_id.
After each task the weights are checkpointed under DDP-style keys and all tasks seen so far are evaluated with greedy
decoding; the prototype banks are saved at the end (vqacl.py:413-426).

Only the data is synthetic (no VQA v2 here): `SyntheticTaskData` produces collate_fn-shaped dicts
(vqa_data_memory.py:365-394) whose question-type / category one-hots follow Question_type.py's tables.

    python examples/vqacl_task_loop.py --tasks 3 --groups 2 --iters 4 --batch_size 32 --layers 2
"""
import argparse
import itertools
import os
import random
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vqacl_b200 as V  # noqa: E402
from vqacl_b200.continual import RehearsalMemory, load_training_state, save_training_state  # noqa: E402

ALL_TASKS = ["q_recognition", "q_location", "q_judge", "q_commonsense", "q_count", "q_action", "q_color", "q_type",
             "q_subcategory", "q_causal"]                                    # Question_type.py:16
N_GROUPS = 5                                                                  # Category_splits G1..G5, Question_type.py:20-24
UNREACHABLE = (0, 12, 26, 29, 30, 45, 66, 68, 69, 71)


def category_splits(n_groups):
    slots = [c for c in range(80) if c not in UNREACHABLE]
    per = (len(slots) + N_GROUPS - 1) // N_GROUPS
    return {f"G{g + 1}": slots[g * per:(g + 1) * per] for g in range(n_groups)}


class SyntheticTaskData:
    """Batches of one (question-type task, category group); `old_tasks` makes a rehearsal loader over earlier tasks."""

    def __init__(self, task_id, cates, n_batches, batch_size, seed, old_tasks=None, vocab=32000):
        self.task_id, self.cates, self.n, self.B, self.seed, self.old, self.vocab = task_id, cates, n_batches, batch_size, seed, old_tasks, vocab

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            g = torch.Generator().manual_seed(self.seed * 1000 + i)
            B = self.B if i + 1 < self.n else max(2, self.B // 2)            # ragged last batch: no drop_last in the reference
            L = int(torch.randint(12, 21, (1,), generator=g))                # text width = batch max <= 20
            T = int(torch.randint(2, 7, (1,), generator=g))
            lens = torch.randint(4, L + 1, (B,), generator=g)
            lens[0] = L
            ids = torch.randint(3, self.vocab, (B, L), generator=g)
            pos = torch.arange(L)[None, :]
            ids = torch.where(pos < lens[:, None] - 1, ids, torch.zeros_like(ids))
            ids[torch.arange(B), lens - 1] = 1
            tl = torch.randint(2, T + 1, (B,), generator=g)
            tl[0] = T
            tgt = torch.randint(3, self.vocab, (B, T), generator=g)
            tgt = torch.where(torch.arange(T)[None, :] < tl[:, None] - 1, tgt, torch.full_like(tgt, -100))
            tgt[torch.arange(B), tl - 1] = 1
            xy = torch.rand(B, 36, 2, generator=g) * 0.7
            boxes = torch.cat([xy, (xy + torch.rand(B, 36, 2, generator=g) * 0.25 + 0.05).clamp(max=1.0)], dim=2)
            cate = torch.zeros(B, 80)
            cate[torch.arange(B), torch.tensor(self.cates)[torch.randint(0, len(self.cates), (B,), generator=g)]] = 1
            ques = torch.zeros(B, 10)
            q = (torch.tensor(self.old)[torch.randint(0, len(self.old), (B,), generator=g)] if self.old
                 else torch.full((B,), self.task_id))
            ques[torch.arange(B), q] = 1
            yield dict(vis_feats=torch.relu(torch.randn(B, 36, 2048, generator=g)), boxes=boxes, input_ids=ids, target_ids=tgt,
                       scores=torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (B,), generator=g)], cate_labels=cate,
                       ques_labels=ques)


def task_records(task_idx, n=200):
    """Stand-in for datasets/vqa/Partition_Q/karpathy_train_<task>.json: training records of one task with an image id each."""
    g = random.Random(4000 + task_idx)
    return [{"img_id": f"img{g.randrange(500)}", "question_id": f"t{task_idx}-{k}", "seed": g.randrange(1 << 30)} for k in range(n)]


IMG_CATE = {f"img{i}": (i * 7) % 90 + 1 for i in range(500)}                  # raw COCO ids 1..90 as in ImgId_cate_map.json (H10)


def run(args, log=print):
    dev = torch.device("cuda", 0)
    torch.manual_seed(args.seed)
    random.seed(args.seed)
    cfg = V.VLT5Config(num_layers=args.layers, num_decoder_layers=args.layers, dropout_rate=args.dropout)
    model = V.VLT5VQA.from_pretrained("t5-base", config=cfg)
    model.resize_token_embeddings(32200)
    model = model.to(dev)
    splits = category_splits(args.groups)
    out_dir = args.output or tempfile.mkdtemp(prefix="vqacl_")
    os.makedirs(out_dir, exist_ok=True)
    history = []
    memory = RehearsalMemory(splits, M=getattr(args, "m_size", 64))
    first_task = 0
    resume = getattr(args, "resume", None)
    if resume:                                                               # true resume: weights + banks + bookkeeping + memory + RNG
        last, _ = load_training_state(resume, model, memory)
        first_task = last + 1
        log(f"resumed after task {last}")
    for task_idx, task in enumerate(ALL_TASKS[:args.tasks]):
        if task_idx < first_task:
            continue
        log(f"======== task {task_idx} {task} ========")
        if args.memory and task_idx > 0:                                     # vqacl.py:169-203
            all_ex, each = memory.grow(task_idx, task_records(task_idx - 1), IMG_CATE)
            log(f"  rehearsal memory: {len(all_ex)} exemplars ({each} per old task)")
        groups = list(splits)
        random.shuffle(groups)                                               # random_dic(Category_splits), vqacl.py:314
        for group in groups:
            train = SyntheticTaskData(task_idx, splits[group], args.iters, args.batch_size, seed=task_idx * 17 + int(group[1:]))
            # rehearsal loader over the exemplar memory (here: synthetic batches seeded by the stored exemplar records)
            mem_loader = (SyntheticTaskData(task_idx, splits[group], max(1, args.iters // 2), args.batch_size,
                                            seed=999 + task_idx + sum(d["seed"] for d in all_ex) % 1000,
                                            old_tasks=list(range(task_idx))) if task_idx > 0 and args.memory else None)
            total = (2 if mem_loader else 1) * len(train) * args.batch_size
            t_total = int(total / args.batch_size) * args.epochs
            optim = V.FusedAdamW(model, lr=args.lr, eps=1e-6, weight_decay=0.01)     # new optimizer per group (vqacl.py:329)
            sched = V.get_constant_schedule_with_warmup(optim, int(t_total * 0.1))
            for epoch in range(args.epochs):
                model.train()
                loader = zip(train, itertools.cycle(mem_loader)) if mem_loader else ((b, None) for b in train)
                for batch, mem_batch in loader:
                    for b in (batch, mem_batch):
                        if b is None:
                            continue
                        res = model.train_step(b, task_idx, args.proto_alpha, args.proto_beta)      # vqacl.py:438
                        res["loss"].backward()
                        optim.step(max_grad_norm=5.0)
                        sched.step()
                        optim.zero_grad()
                        history.append(res["loss"])
        ckpt = os.path.join(out_dir, f"{task}_LAST.pth")
        torch.save({"module." + k: v.cpu() for k, v in model.state_dict().items()}, ckpt)      # trainer_base.py:246-249
        save_training_state(os.path.join(out_dir, f"{task}_STATE.pt"), model, task_idx, memory)  # what a resume needs on top
        # evaluate every task seen so far with greedy decoding (vqacl.py:416-417, 545-579)
        model.eval()
        for t in range(task_idx + 1):
            tb = next(iter(SyntheticTaskData(t, splits["G1"], 1, args.batch_size, seed=5000 + t)))
            out = model.test_step(tb)
            log(f"  test task {t}: generated {tuple(out['token_ids'].shape)} tokens")
    torch.save(model.Q_prototype.cpu(), os.path.join(out_dir, "Q_prototype.pt"))               # vqacl.py:419-423
    torch.save(model.V_prototype.cpu(), os.path.join(out_dir, "V_prototype.pt"))
    losses = torch.stack([h.detach() for h in history]).cpu()
    log(f"steps {len(losses)}  first loss {losses[0]:.3f}  last loss {losses[-1]:.3f}  output {out_dir}")
    return model, losses, out_dir


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=3)
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--proto_alpha", type=float, default=0.5)
    ap.add_argument("--proto_beta", type=float, default=0.3)
    ap.add_argument("--memory", action="store_true", default=True)
    ap.add_argument("--seed", type=int, default=66666)
    ap.add_argument("--output", default=None)
    run(ap.parse_args())


if __name__ == "__main__":
    main()
