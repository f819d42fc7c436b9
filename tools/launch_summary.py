"""Summarise an ncu `gpu__time_duration.sum` launch list (CSV) for one train step: per-kernel totals and per-phase totals.
python tools/launch_summary.py gpurun_out/launches.csv [step_index]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return [(x["Kernel Name"], float(x["Metric Value"].replace(",", "")), x["Grid Size"]) for x in csv.DictReader(lines)]


def short(n):
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("vq::", "")


def main():
    rows = load(sys.argv[1])
    starts = [i for i, (n, _, _) in enumerate(rows) if "keymask" in n]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if len(starts) == 1:
        # a window of ~one step's worth of consecutive launches taken mid-run (--launch-skip/--launch-count): rotate it so
        # that it starts at the step boundary (the steady-state step is cyclic; at most a launch or two are missing)
        step = rows[starts[0]:] + rows[:starts[0]]
        print("(cyclic window: the capture holds one step boundary; rotated to start there)")
    elif len(starts) < k + 2:
        print("need two step starts (keymask_kernel) in the capture; found", len(starts)); return
    else:
        step = rows[starts[k]:starts[k + 1]]
    names = [short(n) for n, _, _ in step]
    tot = sum(t for _, t, _ in step)
    print(f"launches in step: {len(step)}   sum of kernel durations: {tot / 1e6:.3f} ms")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, (_, t, _) in zip(names, step):
        agg[n][0] += 1
        agg[n][1] += t
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}%  x{c:4d}  {n[:100]}")
    def idx(name, nth=0):
        return [i for i, n in enumerate(names) if n == name][nth]
    try:
        marks = [("encoder fwd", 0, idx("proto_means_kernel")), ("SI path", idx("proto_means_kernel"), idx("shift_right_kernel")),
                 ("decoder fwd + LM head + CE", idx("shift_right_kernel"), idx("ce_fwd_kernel") + 1),
                 ("loss tail", idx("ce_fwd_kernel") + 1, idx("ce_bwd_kernel")),
                 ("LM head + decoder bwd", idx("ce_bwd_kernel"), idx("embed_bwd_kernel")),
                 ("cross-KV + encoder bwd + embeddings", idx("embed_bwd_kernel"), idx("sumsq_partial_kernel")),
                 ("clip + AdamW", idx("sumsq_partial_kernel"), len(step))]
        for nm, a, b in marks:
            print(f"{sum(t for _, t, _ in step[a:b]) / 1e3:10.1f} us  {b - a:4d} launches  {nm}")
    except (IndexError, ValueError):
        pass


if __name__ == "__main__":
    main()
