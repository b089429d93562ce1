#!/usr/bin/env python
"""Kernel timeline of the CUDA-graph training step (bench.py's `value` loop): per-kernel GPU time, and the critical-path
split 'a recurrence kernel is running' vs 'no recurrence kernel is running' (everything in the second bucket is exposed
non-recurrent work: GEMMs, attention, encoder, optimizer, launch gaps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from flowtron_b200 import synth
from flowtron_b200.flowtron import Flowtron, FlowtronLoss

cfg = dict(synth.DEFAULT_MODEL_CONFIG)
model = Flowtron(**cfg); model.load_state_dict(synth.synth_params(cfg, 1234), strict=True); model = model.cuda().train()
crit = FlowtronLoss()
opt = torch.optim.RAdam(model.parameters(), lr=1e-3, weight_decay=1e-6, capturable=True, foreach=True)
_ol, _il = synth.ljs_like_lengths(32, 1000, 1234)
L = int(_il.max())
batch = synth.synth_batch(32, 1000, L, cfg, 1234, out_lens=_ol.tolist(), in_lens=_il.tolist(), with_prior=True, logmel_stats=True)
d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}

def step():
    opt.zero_grad(set_to_none=True)
    out = model(d["mel"], d["speaker_ids"], d["text"], d["in_lens"], d["out_lens"], d["attn_prior"])
    nll, gl, _ = crit(out, d["gate_target"], d["in_lens"], d["out_lens"])
    (nll + gl).sum().backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0, foreach=True)
    opt.step()

for _ in range(3): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    step()
g.replay(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay(); g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
def union(evs):
    tot, cur_s, cur_e = 0.0, None, None
    for e in sorted(evs, key=lambda e: e.time_range.start):
        s, en = e.time_range.start, e.time_range.end
        if cur_e is None or s > cur_e:
            if cur_e is not None: tot += cur_e - cur_s
            cur_s, cur_e = s, en
        else:
            cur_e = max(cur_e, en)
    if cur_e is not None: tot += cur_e - cur_s
    return tot
rec = [e for e in ev if "lstm_fwd_kernel" in e.name or "lstm_bwd" in e.name]
busy, rec_busy = union(ev), union(rec)
print(f"wall {(t1 - t0) / 2e3:.2f} ms/step; GPU busy {busy / 2e3:.2f}; idle {(t1 - t0 - busy) / 2e3:.2f}; a recurrence kernel running {rec_busy / 2e3:.2f}; "
      f"exposed non-recurrent {(t1 - t0 - rec_busy) / 2e3:.2f} ms/step; kernels/step {len(ev) // 2}")
# what runs while no recurrence kernel is running
spans = []
cur_s = cur_e = None
for e in sorted(rec, key=lambda e: e.time_range.start):
    s, en = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None: spans.append((cur_s, cur_e))
        cur_s, cur_e = s, en
    else:
        cur_e = max(cur_e, en)
spans.append((cur_s, cur_e))
def exposed(e):
    s, en = e.time_range.start, e.time_range.end
    cov = 0.0
    for a, b in spans:
        lo, hi = max(s, a), min(en, b)
        if hi > lo: cov += hi - lo
    return (en - s) - cov
agg = {}
for e in ev:
    if e in rec: continue
    x = exposed(e)
    if x <= 0: continue
    k = e.name[:60]
    a = agg.setdefault(k, [0.0, 0]); a[0] += x; a[1] += 1
print("kernel time outside recurrence spans (ms/step, launches/step) -- overlapping kernels are each counted:")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {t / 2e3:8.3f} {n // 2:5d}  {k}")
gaps = []
for (a, b), (c, dd) in zip(spans[:-1], spans[1:]):
    gaps.append(c - b)
gaps.sort(reverse=True)
print("largest windows without a recurrence kernel (ms):", [round(x / 1e3, 3) for x in gaps[:24]])
