"""Debug helper: where does a resumed run diverge from the uninterrupted one? (compares the per-task STATE files)"""
import importlib.util, os, sys, tempfile, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_err
spec = importlib.util.spec_from_file_location("loop", os.path.join(ROOT, "examples", "vqacl_task_loop.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
base = dict(groups=2, iters=2, epochs=1, batch_size=8, layers=1, lr=1e-3, dropout=0.0, proto_alpha=0.5, proto_beta=0.3, memory=True, seed=3, m_size=24)
tmp = tempfile.mkdtemp()
full, part = os.path.join(tmp, "full"), os.path.join(tmp, "part")
logs = {"full": [], "part": [], "res": []}
mod.run(types.SimpleNamespace(tasks=3, output=full, **base), log=lambda *a: logs["full"].append(" ".join(map(str, a))))
mod.run(types.SimpleNamespace(tasks=2, output=part, **base), log=lambda *a: logs["part"].append(" ".join(map(str, a))))
mod.run(types.SimpleNamespace(tasks=3, output=part, resume=os.path.join(part, "q_location_STATE.pt"), **base), log=lambda *a: logs["res"].append(" ".join(map(str, a))))
for k, v in logs.items():
    print(k, v)
for name in ("q_location_STATE.pt", "q_judge_STATE.pt"):
    a, b = torch.load(os.path.join(full, name), weights_only=False), torch.load(os.path.join(part, name), weights_only=False)
    print(name, "task", a["task_idx"], b["task_idx"], "cur", a["Q_task_cur_proto"], b["Q_task_cur_proto"], "mem", a["Q_task_mem_proto"], b["Q_task_mem_proto"])
    print("  rng python equal", a["rng"]["python"] == b["rng"]["python"], "torch equal", torch.equal(a["rng"]["torch"], b["rng"]["torch"]),
          "memory equal", a["memory"] == b["memory"], "step_seed", a["step_seed"], b["step_seed"])
    for k in ("Q_prototype", "V_prototype", "Q_prototype_num", "V_prototype_num"):
        print("  ", k, rel_err(a[k], b[k]))
    w = max((rel_err(a["model"][k], b["model"][k]), k) for k in a["model"])
    print("  worst weight", w)
    print("  Q rows rel", [(round(rel_err(a["Q_prototype"][i], b["Q_prototype"][i]), 4) if a["Q_prototype"][i].abs().max() > 0 else 0) for i in range(4)])
