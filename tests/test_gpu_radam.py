"""GPU: fused clip + RAdam (csrc/optim.cu through the C ABI) against the reference optimizer's recorded trajectory
(tests/golden/radam.npz) and against the CPU oracle on a real model's gradients.  fp32: rtol 1e-5 / atol 1e-7 (FMA
contraction and reduction order are the only differences)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _golden():
    return np.load(os.path.join(GOLDEN, "radam.npz"))


@pytest.mark.parametrize("mode", ["flat", "fresh"])
def test_radam_trajectory_matches_reference(mode):
    from flowtron_b200.radam import RAdam
    g = _golden()
    ps = [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"]).cuda()) for i in range(3)]
    opt = RAdam(ps, lr=1e-3, weight_decay=1e-6)
    for s in range(int(g["n_steps"])):
        for i, p in enumerate(ps):
            gi = torch.from_numpy(g[f"g{s}_{i}"]).cuda()
            if mode == "fresh":
                p.grad = gi
            else:
                p.grad.copy_(gi)
        total = opt.clip_grad_norm_(float(g["max_norm"]))
        opt.step()
        assert abs(float(total) - float(g[f"norm{s}"])) <= 1e-5 * float(g[f"norm{s}"])
        for i, p in enumerate(ps):
            torch.testing.assert_close(p.detach().cpu(), torch.from_numpy(g[f"p{s + 1}_{i}"]), rtol=1e-5, atol=1e-7)
        if mode == "fresh":
            opt.zero_grad(set_to_none=True)
    for i, p in enumerate(ps):
        torch.testing.assert_close(opt.state[p]["exp_avg"].cpu(), torch.from_numpy(g[f"m_{i}"]), rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(opt.state[p]["exp_avg_sq"].cpu(), torch.from_numpy(g[f"v_{i}"]), rtol=1e-5, atol=1e-10)


def test_radam_on_flowtron_model_matches_oracle():
    """Re-homing the parameters into the flat buffer keeps the model working; two clipped steps on real gradients equal
    the CPU oracle parameter by parameter; everything is ONE update launch."""
    from flowtron_b200 import _lib, synth
    from flowtron_b200.flowtron import Flowtron, FlowtronLoss
    from flowtron_b200.radam import RAdam
    from oracle import radam_oracle as O
    cfg = dict(synth.DEFAULT_MODEL_CONFIG, n_flows=2)
    model = Flowtron(**cfg)
    model.load_state_dict(synth.synth_params(cfg, 11), strict=True)
    model = model.cuda().train()
    model.encoder.p_dropout = 0.0
    names = [n for n, _ in model.named_parameters()]
    before = {n: p.detach().cpu().clone() for n, p in model.named_parameters()}
    opt = RAdam(model.parameters(), lr=1e-3, weight_decay=1e-6)
    for n, p in model.named_parameters():
        assert torch.equal(p.detach().cpu(), before[n]) and p.data_ptr() % 256 == 0
    crit = FlowtronLoss(sigma=1.0, gate_loss=True, use_ctc_loss=False)
    b = synth.synth_batch(3, 40, 12, cfg, 4, out_lens=[40, 31, 17], in_lens=[12, 9, 5])
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    ref_p = dict(before)
    ref_m = {n: torch.zeros_like(v) for n, v in before.items()}
    ref_v = {n: torch.zeros_like(v) for n, v in before.items()}
    for step in (1, 2):
        opt.zero_grad()
        out = model(d["mel"], d["speaker_ids"], d["text"], d["in_lens"], d["out_lens"], None)
        nll, gl, _ = crit(out, d["gate_target"], d["in_lens"], d["out_lens"])
        (nll + gl).sum().backward()
        grads = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}
        assert len(grads) == len(names)
        _lib.reset_launch_count()
        total = opt.clip_grad_norm_(1.0)
        opt.step()
        assert _lib.launch_count() == 3                   # partial sums, finalize, update
        t_ref, coef = O.clip_coef([grads[n] for n in names], 1.0)
        assert abs(float(total) - float(t_ref)) <= 1e-5 * float(t_ref)
        for n in names:
            ref_p[n], ref_m[n], ref_v[n] = O.radam_step(ref_p[n], grads[n] * coef, ref_m[n], ref_v[n], step, lr=1e-3,
                                                         weight_decay=1e-6)
        for n, p in model.named_parameters():
            torch.testing.assert_close(p.detach().cpu(), ref_p[n], rtol=1e-5, atol=1e-7, msg=lambda m, n=n: f"{n}: {m}")
    assert torch.isfinite(nll).all()


def test_radam_capturable_replayed_from_a_cuda_graph_matches_reference():
    """capturable=True: clip + step captured ONCE in a CUDA graph and replayed for every step of the reference trajectory
    (steps 1-5 take the N_sma < 5 branch, 6+ the adaptive one: the kernel must derive both from the device step count)."""
    from flowtron_b200.radam import RAdam
    g = _golden()
    ps = [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"]).cuda()) for i in range(3)]
    opt = RAdam(ps, lr=1e-3, weight_decay=1e-6, capturable=True)
    norm_out = torch.zeros((), device="cuda")

    def body():
        total = opt.clip_grad_norm_(float(g["max_norm"]))
        opt.step()
        norm_out.copy_(total)

    s_ = torch.cuda.Stream()
    s_.wait_stream(torch.cuda.current_stream())
    graph = None
    for s in range(int(g["n_steps"])):
        for i, p in enumerate(ps):
            p.grad.copy_(torch.from_numpy(g[f"g{s}_{i}"]).cuda())
        if s == 0:                                   # eager first step (warm-up), then capture, then replays only
            body()
        elif graph is None:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            graph.replay()
        else:
            graph.replay()
        torch.cuda.synchronize()
        assert abs(float(norm_out) - float(g[f"norm{s}"])) <= 1e-5 * float(g[f"norm{s}"])
        for i, p in enumerate(ps):
            torch.testing.assert_close(p.detach().cpu(), torch.from_numpy(g[f"p{s + 1}_{i}"]), rtol=1e-5, atol=1e-7)
    sd = opt.state_dict()
    assert all(int(v["step"]) == int(g["n_steps"]) for v in sd["state"].values())


def test_radam_survives_parameter_rehoming_by_cudnn():
    """nn.LSTM.flatten_parameters() (called by cuDNN paths, e.g. Encoder.infer in the reference) moves weights
    out of the flat buffer; the optimizer must re-home them instead of updating a dead copy."""
    from flowtron_b200.radam import RAdam
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(16, 8, 1, batch_first=True, bidirectional=True).cuda()
    opt = RAdam(lstm.parameters(), lr=1e-2)
    x = torch.randn(4, 5, 16, device="cuda")

    def one_step():
        opt.zero_grad()
        lstm(x)[0].pow(2).sum().backward()
        opt.step()
    one_step()
    w0 = lstm.weight_hh_l0.detach().clone()
    lstm.flatten_parameters()                          # cuDNN re-packs: p.data now lives in a fresh buffer
    one_step()
    assert not torch.equal(lstm.weight_hh_l0.detach(), w0), "the live weights must keep training"
    base = opt._flats[0].p.data_ptr()
    assert all(p.data_ptr() == base + 4 * off for p, off in zip(opt._flats[0].params, opt._flats[0].offsets))
