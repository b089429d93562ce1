#!/usr/bin/env python
"""Per-phase timing of the persistent inference kernel (clock64 stamps of CTA 0) -> where a frame's time goes."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowtron_b200 import _lib, synth
from flowtron_b200.flowtron import Flowtron

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T, L = 200, 100
cfg = dict(synth.DEFAULT_MODEL_CONFIG, use_gate_layer=False)
m = Flowtron(**cfg); m.load_state_dict(synth.synth_params(cfg, 1), strict=True); m = m.cuda().eval()
g = torch.Generator().manual_seed(0)
res = (torch.randn(B, 80, T, generator=g) * 0.5).cuda(); text = torch.randint(0, 185, (B, L), generator=g).cuda()
spk = torch.zeros(B, dtype=torch.long, device="cuda")
trace = torch.zeros(T, 32, dtype=torch.int64, device="cuda")
Lb = _lib.lib(); Lb.ft_debug_set_infer_trace.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    m.infer(res, spk, text)
    Lb.ft_debug_set_infer_trace(ctypes.c_void_p(trace.data_ptr()))
    m.flows[0].infer(res.permute(2, 0, 1).contiguous(), torch.randn(L, B, 640, device="cuda"), None)
    torch.cuda.synchronize()
    Lb.ft_debug_set_infer_trace(None)
tr = trace.cpu().double()[20:180]
names = {0: "frame start", 1: "P1 inputs staged", 2: "P1 matvec+cell", 4: "P2 input staged", 5: "P2 matvec (q)", 6: "P3a scores",
         7: "P3b softmax+ctx+gate", 9: "P4 inputs staged", 10: "P4 matvec+cell", 12: "P5 inputs staged", 13: "P5 matvec+cell",
         15: "P6 dense", 17: "P7 dense", 18: "P8 input staged", 19: "P8 conv+out"}
period = (tr[1:, 0] - tr[:-1, 0]).mean().item()
print(f"B={B}: frame period {period:.0f} clk = {period / 1965:.2f} us")
prev = 0.0
for k in sorted(names):
    v = (tr[:, k] - tr[:, 0]).mean().item()
    print(f"  {names[k]:22s} +{v / 1965:7.2f} us   (d {(v - prev) / 1965:5.2f})")
    prev = v
