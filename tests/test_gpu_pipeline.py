"""GPU: the chunk-pipelined forward and BPTT of lstm layers 0/1 (FT_PIPE_FWD=1 / FT_PIPE_BWD=1, DESIGN.md 4.2: two 64-CTA
recurrences one chunk apart) must reproduce the one-launch-per-layer schedule bit for bit -- same kernel, same arithmetic,
different step ranges per launch.  The switch is read once per process by the library, so each run is a child process."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

CHILD = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])
from flowtron_b200 import synth
from flowtron_b200.flowtron import Flowtron
cfg = dict(synth.DEFAULT_MODEL_CONFIG, n_flows=2)
model = Flowtron(**cfg); model.load_state_dict(synth.synth_params(cfg, 5), strict=True); model = model.cuda()
b = synth.synth_batch(5, 77, 12, cfg, 9, with_prior=True)
d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
from flowtron_b200.flowtron import FlowtronLoss
model.train(); model.encoder.p_dropout = 0.0
torch.backends.cudnn.allow_tf32 = False
out = model(d["mel"], d["speaker_ids"], d["text"], d["in_lens"], d["out_lens"], d["attn_prior"])
nll, gl, _ = FlowtronLoss()(out, d["gate_target"], d["in_lens"], d["out_lens"])
(nll + gl).sum().backward()
torch.cuda.synchronize()
torch.save({"z": out[0].detach().cpu(), "log_s": [x.detach().cpu() for x in out[1]], "gate": out[2].detach().cpu(),
            "grads": {n: p.grad.cpu() for n, p in model.named_parameters()}}, sys.argv[2])
'''


def _run(tmp_path, name, env):
    out = tmp_path / f"{name}.pt"
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, str(out)], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


def test_pipelined_layers_equal_default_schedule(tmp_path):
    a = _run(tmp_path, "default", {"FT_PIPE_FWD": "0", "FT_PIPE_BWD": "0"})
    b = _run(tmp_path, "pipe", {"FT_PIPE_FWD": "1", "FT_PIPE_BWD": "1", "FT_PIPE_CHUNK": "16"})   # 77 steps -> 5 chunks, ragged tail
    assert torch.equal(a["z"], b["z"]) and torch.equal(a["gate"], b["gate"])
    for x, y in zip(a["log_s"], b["log_s"]):
        assert torch.equal(x, y)
    gmax = max(v.norm().item() for v in a["grads"].values())
    for k in a["grads"]:
        d = (a["grads"][k] - b["grads"][k]).norm().item()
        assert d <= 1e-3 * (a["grads"][k].norm().item() + 1e-4 * gmax), (k, d)     # fp32 atomics / split-K order only
