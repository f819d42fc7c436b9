"""In-step kernel timeline of the bench step (B=320) from CUPTI (torch.profiler): unlike ncu's serialised per-launch
durations this shows what each kernel costs on the real, overlapped timeline, per stream, and how much of the step the
GPU sits idle between dependent launches. Writes gpurun_out/step_trace.json (chrome trace) and prints a summary.

    python tools/step_trace.py [B] [steps]
"""
import json, os, sys, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch
from torch.profiler import profile, ProfilerActivity
import vqacl_b200 as V
import vlt5_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
m = V.VLT5VQA(V.VLT5Config(vocab_size=32200, dropout_rate=0.1)).to("cuda")
m.train()
opt = V.FusedAdamW(m)
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_batch(B, task_id=3).items()}


def step():
    out = m.train_step(batch, 3, 0.5, 0.3)
    out["loss"].backward()
    opt.step(max_grad_norm=5.0)
    for p in m.parameters():
        p.grad = None


for _ in range(6):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/step_trace.json"
prof.export_chrome_trace(path)

ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
span = (t1 - t0) / steps
per = collections.defaultdict(lambda: [0, 0.0])
streams = collections.defaultdict(float)
for e in ev:
    n = e["name"].split("(")[0][:70]
    per[n][0] += 1
    per[n][1] += e["dur"]
    streams[e["args"].get("stream")] += e["dur"]
# union of busy intervals over all streams
busy, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, f = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, f
    else:
        cur_e = max(cur_e, f)
busy += cur_e - cur_s
print(f"B={B} steps={steps}: span/step {span / 1e3:.3f} ms, GPU busy (any stream) {busy / steps / 1e3:.3f} ms, idle {100 * (1 - busy / (t1 - t0)):.1f} %")
for s, d in sorted(streams.items(), key=lambda kv: -kv[1]):
    print(f"  stream {s}: {d / steps / 1e3:.3f} ms of kernels per step")
print("per step: us  count  avg_us  kernel")
for n, (c, d) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{d / steps:10.1f} {c / steps:6.0f} {d / c:8.1f}  {n}")

# timeline excerpt of the last profiled step: N kernels starting at a named one (default: the LM-head backward, i.e. the decoder
# backward chain):  python tools/step_trace.py 320 3 ce_bwd_kernel 120
if len(sys.argv) > 3:
    start_name, count = sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 100
    idx = [i for i, e in enumerate(ev) if start_name in e["name"]]
    if idx:
        i0 = idx[-1]
        base = ev[i0]["ts"]
        print(f"--- timeline from the last {start_name} (us: start, end, duration | stream | kernel)")
        prev_end = {}
        for e in ev[i0:i0 + count]:
            st = e["args"].get("stream")
            gap = e["ts"] - prev_end.get(st, e["ts"])
            prev_end[st] = e["ts"] + e["dur"]
            print(f"{e['ts'] - base:9.1f} {e['ts'] + e['dur'] - base:9.1f} {e['dur']:7.1f} gap {gap:6.1f} | {st} | {e['name'].split('(')[0][:60]}")
