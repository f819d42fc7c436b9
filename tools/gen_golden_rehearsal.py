"""Generate tests/golden/rehearsal_memory.json by EXECUTING the reference's own rehearsal-memory block
(VL-T5/src/vqacl.py:169-203, the body of `if args.memory:` inside Trainer.train's task loop) on synthetic partition files.

The block is lifted as text, dedented and run unmodified for task_idx = 1..4 in a namespace that provides exactly the names
it touches (`self.M`, `self.task_list`, `self.Examplar_set`, `Category_splits`, `ImgId_cate_map`, `json`, `random`,
`latest_task_idx`); the partition JSONs it opens by relative path are written to a temporary working directory.
    python tools/gen_golden_rehearsal.py [/root/reference]
"""
import json
import os
import random
import sys
import tempfile
import textwrap
import types

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rehearsal_memory.json")


def main():
    lines = open(os.path.join(REF, "VL-T5", "src", "vqacl.py")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "if task_idx != latest_task_idx + 1:")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip().startswith("print(\"# The size of the cate Memory:\""))
    block = textwrap.dedent("\n".join(lines[i0:i1 + 1]))
    tasks = ["q_recognition", "q_location", "q_judge", "q_commonsense", "q_count"]
    splits = {"G1": [1, 2, 3, 4], "G2": [5, 6, 7], "G3": [8, 9, 10, 11, 13]}
    rng = random.Random(7)
    img_cate = {f"img{i}": rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 80, 85]) for i in range(400)}   # ids >= 80: in no group (H10)
    partitions = {t: [{"img_id": f"img{rng.randrange(440)}", "question_id": f"{t}-{k}"} for k in range(rng.randrange(60, 140))] for t in tasks}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "datasets", "vqa", "Partition_Q"))
    for t, d in partitions.items():
        json.dump(d, open(os.path.join(tmp, "datasets", "vqa", "Partition_Q", f"karpathy_train_{t}.json"), "w"))
    os.chdir(tmp)
    try:
        self = types.SimpleNamespace(M=100, task_list=tasks, Examplar_set={"G1": [], "G2": [], "G3": []})
        steps = []
        for task_idx in range(1, 5):
            random.seed(1000 + task_idx)
            ns = dict(self=self, task_idx=task_idx, latest_task_idx=-1, json=json, random=random, Category_splits=splits,
                      ImgId_cate_map=img_cate, print=lambda *a, **k: None)
            exec(block, ns)
            steps.append(dict(task_idx=task_idx, seed=1000 + task_idx, each_memory=ns["each_memory"],
                              all_examplar=[d["question_id"] for d in ns["All_examplar"]],
                              examplar_set={g: [[d["question_id"] for d in ts] for ts in v] for g, v in self.Examplar_set.items()}))
    finally:
        os.chdir(cwd)
    json.dump(dict(M=100, tasks=tasks, splits=splits, img_cate=img_cate, partitions=partitions, steps=steps), open(OUT, "w"))
    print(OUT, os.path.getsize(OUT), [len(s["all_examplar"]) for s in steps])


if __name__ == "__main__":
    main()
