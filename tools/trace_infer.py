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
names = ["frame start", "P1 staged", "P1 matvec+cell", "P1 barrier", "P2 staged", "P2 matvec", "P2 barrier", "P3 attention", "P3 barrier",
         "P4 staged", "P4 matvec+cell", "P4 barrier", "P5 staged", "P5 matvec+cell", "P5 barrier", "P6 dense", "P6 barrier", "P7 dense",
         "P7 barrier", "P8 conv+out", "P8 barrier"]
period = (tr[1:, 0] - tr[:-1, 0]).mean().item()
print(f"B={B}: frame period {period:.0f} clk = {period / 1965:.2f} us")
prev = 0.0
for k, n in enumerate(names):
    v = (tr[:, k] - tr[:, 0]).mean().item()
    print(f"  {n:18s} +{v / 1965:7.2f} us   (d {(v - prev) / 1965:5.2f})")
    prev = v
