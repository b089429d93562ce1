#!/usr/bin/env python
"""GPU-idle accounting of one training step with torch.profiler (kernel timeline): busy time vs wall, largest gaps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from flowtron_b200 import synth
from flowtron_b200.flowtron import Flowtron, FlowtronLoss

cfg = dict(synth.DEFAULT_MODEL_CONFIG)
model = Flowtron(**cfg); model.load_state_dict(synth.synth_params(cfg, 1234), strict=True); model = model.cuda().train()
crit = FlowtronLoss()
FUSED = os.environ.get("FT_FUSED_OPT", "0") != "0"        # same switch as bench.py: profile the fused optimizer's stall
if FUSED:
    from flowtron_b200.radam import RAdam
    opt = RAdam(model.parameters(), lr=1e-3, weight_decay=1e-6)
else:
    opt = torch.optim.RAdam(model.parameters(), lr=1e-3, weight_decay=1e-6)
_ol, _il = synth.ljs_like_lengths(32, 1000, 1234)
L = int(_il.max())
batch = synth.synth_batch(32, 1000, L, cfg, 1234, out_lens=_ol.tolist(), in_lens=_il.tolist(), with_prior=True, logmel_stats=True)
d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}

def step():
    if FUSED:
        opt.zero_grad()
    else:
        opt.zero_grad(set_to_none=True)
    out = model(d["mel"], d["speaker_ids"], d["text"], d["in_lens"], d["out_lens"], d["attn_prior"])
    nll, gl, _ = crit(out, d["gate_target"], d["in_lens"], d["out_lens"])
    (nll + gl).sum().backward()
    if FUSED:
        opt.clip_grad_norm_(1.0)
    else:
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()

for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, cur_end, gaps, prev = 0.0, t0, [], ev[0]
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, prev.name[:40], e.name[:40]))
        busy += en - s
        cur_end, prev = en, e
    elif en > cur_end:
        busy += en - cur_end
        cur_end, prev = en, e
print(f"wall {(t1 - t0) / 1e3:.2f} ms for 2 steps, GPU busy (union of kernels) {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms, kernels {len(ev)}")
gaps.sort(reverse=True)
print("largest gaps (us, prev kernel -> next kernel):")
for g, a, b in gaps[:25]:
    print(f"  {g:8.1f}  {a}  ->  {b}")
for thr in (5, 20, 100):
    print(f"gaps > {thr} us: sum {sum(g for g, _, _ in gaps if g > thr) / 1e3:.2f} ms, count {sum(1 for g, _, _ in gaps if g > thr)}")
print(f"gaps <= 5 us: sum {sum(g for g, _, _ in gaps if g <= 5) / 1e3:.2f} ms, count {sum(1 for g, _, _ in gaps if g <= 5)}")
agg = {}
for e in ev:
    k = e.name[:48]
    a = agg.setdefault(k, [0.0, 0]); a[0] += e.time_range.end - e.time_range.start; a[1] += 1
print("per-kernel GPU time over 2 steps (ms, launches):")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"  {t / 1e3:8.3f} {n:5d}  {k}")
import time
torch.cuda.synchronize(); c0 = time.perf_counter()
for _ in range(3): step()
c1 = time.perf_counter(); torch.cuda.synchronize(); c2 = time.perf_counter()
print(f"host issue time per step {(c1 - c0) / 3 * 1e3:.1f} ms ; wall per step {(c2 - c0) / 3 * 1e3:.1f} ms")
