"""GPU: batched attention-CTC kernel (csrc/ctc.cu) vs torch's nn.CTCLoss restatement of AttentionCTCLoss
(oracle.flowtron_oracle.attention_ctc_loss, pinned to the reference fixture in tests/test_oracle_golden.py), and the
whole use_ctc_loss training branch -- CTC gradients re-enter the flows through ft_ar_step_bwd's d_attn_logprob, for a
forward and a back step, with the prior on (that path had no test) -- vs the reference fixture train_f2ctc."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, record_parity
from oracle import flowtron_oracle as O
from flowtron_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("time_reversed", [False, True])
def test_ctc_kernel_matches_torch_ctc(time_reversed):
    from flowtron_b200 import _lib
    from flowtron_b200.flowtron import AttentionCTCLoss
    g = torch.Generator().manual_seed(5)
    B, T, L = 6, 70, 19
    in_lens = torch.tensor([19, 12, 1, 7, 19, 30 - 11])
    out_lens = torch.tensor([70, 33, 5, 6, 19, 41])               # row 3: fewer frames than tokens -> inf -> zero_infinity
    lp_nat = (torch.randn(B, T, L, generator=g) * 2.0 - 3.0)
    ref_in = lp_nat.clone().requires_grad_(True)
    ref = O.attention_ctc_loss(ref_in, in_lens, out_lens, -1.0)
    ref.backward()
    if time_reversed:                                              # present the same data in a back step's flipped time
        idx = O.back_step_index(out_lens, T)                       # natural t -> row
        lp_dev = torch.zeros_like(lp_nat)
        for b in range(B):
            lp_dev[b][idx[:, b]] = lp_nat[b]
    else:
        lp_dev = lp_nat
    x = lp_dev.cuda().requires_grad_(True)
    loss = AttentionCTCLoss(blank_logprob=-1)(None, in_lens.cuda(), out_lens.cuda(), x[:, None], time_reversed=time_reversed)
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.device_status() == 0
    gx = x.grad.cpu()
    if time_reversed:
        idx = O.back_step_index(out_lens, T)
        gx = torch.stack([gx[b][idx[:, b]] for b in range(B)])
    e_loss = abs(float(loss) - float(ref)) / abs(float(ref))
    e_grad = (gx - ref_in.grad).abs().max().item() / ref_in.grad.abs().max().item()
    record_parity(f"ctc_kernel_rev{int(time_reversed)}", {"loss": e_loss, "grad": e_grad})
    assert e_loss <= 1e-5 and e_grad <= 1e-4, (e_loss, e_grad)


def test_training_with_ctc_loss_matches_reference():
    from flowtron_b200 import _lib
    from flowtron_b200.flowtron import Flowtron, FlowtronLoss
    gold = dict(np.load(os.path.join(GOLDEN, "train_f2ctc.npz")))
    n_flows, B, T, L = (int(gold[k]) for k in ("cfg_n_flows", "B", "T", "L"))
    w = float(gold["ctc_weight"])
    cfg = dict(synth.DEFAULT_MODEL_CONFIG, n_flows=n_flows)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = Flowtron(**cfg)
    model.load_state_dict(synth.synth_params(cfg, int(gold["seed"])), strict=True)
    model = model.cuda().train()
    model.encoder.p_dropout = 0.0
    batch = synth.synth_batch(B, T, L, cfg, int(gold["seed"]), out_lens=[80, 47, 66], with_prior=True)
    cu = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    out = model(cu["mel"], cu["speaker_ids"], cu["text"], cu["in_lens"], cu["out_lens"], cu["attn_prior"])
    crit = FlowtronLoss(sigma=1.0, gate_loss=True, use_ctc_loss=True, ctc_loss_weight=w)
    nll, gl, ctc = crit(out, cu["gate_target"], cu["in_lens"], cu["out_lens"])
    (nll + gl + ctc * w).sum().backward()
    torch.cuda.synchronize()
    assert _lib.device_status() == 0
    errs = {"nll": abs(float(nll) - float(gold["nll"])) / abs(float(gold["nll"])),
            "gate_loss": abs(float(gl) - float(gold["gate_loss"])) / abs(float(gold["gate_loss"])),
            "loss_ctc": abs(float(ctc) - float(gold["loss_ctc"])) / abs(float(gold["loss_ctc"]))}
    # full gradients vs the CPU oracle (pinned to this fixture's norms on the CPU side)
    op = {k: v.clone().requires_grad_(True) for k, v in synth.synth_params(cfg, int(gold["seed"])).items()}
    oout = O.flowtron_forward(op, batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                              batch["attn_prior"], fast=True)
    onll, ogl = O.flowtron_loss(oout, batch["gate_target"], batch["in_lens"], batch["out_lens"])
    octc = O.flowtron_ctc_loss(oout, batch["in_lens"], batch["out_lens"])
    (onll + ogl + w * octc).sum().backward()
    gmax = max(float(gold[f"gnorm::{n}"]) for n, _ in model.named_parameters())
    rows = []
    for name, p in model.named_parameters():
        g = p.grad.detach().reshape(-1).double().cpu()
        r = op[name].grad.reshape(-1).double()
        gn = float(gold[f"gnorm::{name}"])
        rows.append((name, abs(float(g.norm()) - gn) / (gn + 1e-4 * gmax), float((g - r).norm()) / (float(r.norm()) + 1e-4 * gmax)))
    rows.sort(key=lambda r: -r[2])
    errs["grad_norm_worst"] = max(r[1] for r in rows)
    errs["grad_rel_l2_worst"] = rows[0][2]
    print("ctc training errors:", {k: f"{v:.2e}" for k, v in errs.items()}, rows[:4])
    record_parity("train_f2ctc", errs)
    assert errs["nll"] <= 1e-3 and errs["gate_loss"] <= 1e-3 and errs["loss_ctc"] <= 1e-3, errs
    assert errs["grad_norm_worst"] <= 1e-2 and errs["grad_rel_l2_worst"] <= 4e-2, rows[:6]
