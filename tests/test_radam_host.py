"""CPU: host logic of flowtron_b200.radam.RAdam (flat layout, gradient runs, clip, state, checkpoints) with the three
native entry points replaced by CPU stand-ins that follow oracle.radam_oracle -- the CUDA kernels themselves are
checked in tests/test_gpu_radam.py.  Trajectories are pinned to the reference's (tests/golden/radam.npz)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from flowtron_b200 import _lib
from flowtron_b200 import radam as R


def _alias(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr)))


@pytest.fixture
def cpu_kernels(monkeypatch):
    calls = {"radam": [], "sumsq": 0}

    def radam_step_raw(p_ptr, g_ptr, m_ptr, v_ptr, n, b1, b2, eps, wd_lr, step_size, use_denom, coef_ptr=0):
        p, g, m, v = _alias(p_ptr, n), _alias(g_ptr, n), _alias(m_ptr, n), _alias(v_ptr, n)
        gg = g * (_alias(coef_ptr, 1)[0] if coef_ptr else 1.0)
        v.mul_(b2).add_((1 - b2) * gg * gg)
        m.mul_(b1).add_((1 - b1) * gg)
        if wd_lr:
            p.add_(-wd_lr * p)
        p.add_(-step_size * (m / (v.sqrt() + eps)) if use_denom else -step_size * m)
        calls["radam"].append(n)

    def sumsq_partials_raw(x_ptr, n, part_ptr):
        part = _alias(part_ptr, _lib.SUMSQ_PARTIALS)
        part.zero_()
        part[0] = float((_alias(x_ptr, n).double() ** 2).sum())
        calls["sumsq"] += 1

    def clip_coef(partials, n_partials, max_norm, norm_coef):
        norm = float(partials[:n_partials].double().sum().sqrt())
        norm_coef[0] = norm
        norm_coef[1] = min(1.0, max_norm / (norm + 1e-6))

    monkeypatch.setattr(_lib, "radam_step_raw", radam_step_raw)
    monkeypatch.setattr(_lib, "sumsq_partials_raw", sumsq_partials_raw)
    monkeypatch.setattr(_lib, "clip_coef", clip_coef)
    monkeypatch.setattr(_lib, "lib", lambda: None)
    monkeypatch.setattr(R.RAdam, "_require_cuda", staticmethod(lambda p: None))
    return calls


def _golden():
    return np.load(os.path.join(GOLDEN, "radam.npz"))


def _params(g):
    return [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"]).clone()) for i in range(3)]


def test_requires_cuda_parameters():
    with pytest.raises(_lib.FlowtronB200Error):
        R.RAdam([torch.nn.Parameter(torch.zeros(3))])


@pytest.mark.parametrize("mode", ["flat", "fresh", "bucket"])
def test_trajectory_matches_reference(cpu_kernels, mode):
    g = _golden()
    ps = _params(g)
    values_before = [p.detach().clone() for p in ps]
    opt = R.RAdam(ps, lr=1e-3, weight_decay=1e-6)
    for p, v0 in zip(ps, values_before):                       # re-homing keeps the values
        assert torch.equal(p.detach(), v0)
    bucket = None
    if mode == "bucket":                                        # all-reduce bucket: separate flat buffer, same packing
        offs, tot = R.flat_offsets(ps)
        bucket = torch.zeros(tot)
        for p, off in zip(ps, offs):
            p.grad = bucket[off: off + p.numel()].view_as(p)
    for s in range(int(g["n_steps"])):
        for i, p in enumerate(ps):
            gi = torch.from_numpy(g[f"g{s}_{i}"])
            if mode == "fresh":
                p.grad = gi.clone()
            else:
                p.grad.copy_(gi)
        cpu_kernels["radam"].clear()
        total = opt.clip_grad_norm_(float(g["max_norm"]))
        assert abs(float(total) - float(g[f"norm{s}"])) <= 2e-6 * float(total)
        opt.step()
        if mode == "fresh":
            assert cpu_kernels["radam"] == [35, 33, 1]          # one launch per tensor
        else:
            assert cpu_kernels["radam"] == [64 + 64 + 1]        # one launch: two padded spans + the last tensor
        for i, p in enumerate(ps):
            torch.testing.assert_close(p.detach(), torch.from_numpy(g[f"p{s + 1}_{i}"]), rtol=3e-6, atol=1e-7)
        if mode == "fresh":
            opt.zero_grad(set_to_none=True)
            assert all(p.grad is None for p in ps)
    for i, p in enumerate(ps):
        st = opt.state[p]
        assert st["step"] == int(g["n_steps"])
        torch.testing.assert_close(st["exp_avg"], torch.from_numpy(g[f"m_{i}"]), rtol=3e-6, atol=1e-8)
        torch.testing.assert_close(st["exp_avg_sq"], torch.from_numpy(g[f"v_{i}"]), rtol=3e-6, atol=1e-10)
    sd = opt.state_dict()                                       # reference layout (train.py:131-139)
    assert sorted(sd["state"][0].keys()) == ["exp_avg", "exp_avg_sq", "step"]
    assert set(sd["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "params"}


def test_zero_grad_keeps_flat_buffer_and_skipped_params_lag(cpu_kernels):
    from oracle import radam_oracle as O
    g = _golden()
    ps = _params(g)
    opt = R.RAdam(ps, lr=1e-3)
    ptrs = [p.grad.data_ptr() for p in ps]
    for p in ps:
        p.grad.fill_(1.0)
    opt.zero_grad()
    assert [p.grad.data_ptr() for p in ps] == ptrs and all(float(p.grad.abs().sum()) == 0 for p in ps)
    # parameter 1 gets no gradient on the first two steps: its step counter (and bias correction) lags (radam.py:53-54, 81)
    ref = [p.detach().clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    steps = [0, 0, 0]
    gen = torch.Generator().manual_seed(3)
    for s in range(7):
        grads = [torch.randn(p.shape, generator=gen) for p in ps]
        for i, p in enumerate(ps):
            p.grad = None if (i == 1 and s < 2) else grads[i].clone()
        opt.step()
        for i in range(3):
            if i == 1 and s < 2:
                continue
            steps[i] += 1
            ref[i], m[i], v[i] = O.radam_step(ref[i], grads[i], m[i], v[i], steps[i], lr=1e-3)
        for i, p in enumerate(ps):
            torch.testing.assert_close(p.detach(), ref[i], rtol=3e-6, atol=1e-7)
    assert [opt.state[p]["step"] for p in ps] == [7, 5, 7]


def test_checkpoint_round_trip(cpu_kernels):
    g = _golden()
    gen = torch.Generator().manual_seed(5)
    a = _params(g)
    oa = R.RAdam(a, lr=1e-3, weight_decay=1e-6)
    grads = [[torch.randn(p.shape, generator=gen) for p in a] for _ in range(6)]
    for s in range(3):
        for p, gr in zip(a, grads[s]):
            p.grad.copy_(gr)
        oa.step()
    sd = oa.state_dict()
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    ob = R.RAdam(b, lr=1e-3, weight_decay=1e-6)
    ob.load_state_dict(sd)
    assert ob.state[b[0]]["exp_avg"].data_ptr() == ob._flats[0].m.data_ptr()          # still views of the flat buffers
    for s in range(3, 6):
        for p, q, gr in zip(a, b, grads[s]):
            p.grad.copy_(gr)
            q.grad.copy_(gr)
        oa.step()
        ob.step()
    for p, q in zip(a, b):
        assert torch.equal(p.detach(), q.detach())


def test_allreduce_buckets_take_over_gradients(cpu_kernels):
    """train.py:231-252 order: optimizer first, then distributed.apply_gradient_allreduce re-points .grad into its
    buckets.  The optimizer follows the gradients (one launch per bucket run) and frees its own flat gradient buffer."""
    from flowtron_b200.distributed import _Bucket, _default_buckets

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embedding = torch.nn.Embedding(5, 3)
            self.flows = torch.nn.ModuleList([torch.nn.Linear(3, 7), torch.nn.Linear(7, 2)])
            self.encoder = torch.nn.Linear(3, 3)

    torch.manual_seed(0)
    a, b = Tiny(), Tiny()
    b.load_state_dict(a.state_dict())
    oa, ob = R.RAdam(a.parameters(), lr=1e-3), R.RAdam(b.parameters(), lr=1e-3)
    buckets = [_Bucket(ps) for ps in _default_buckets(b)]
    assert len(buckets) == 3
    gen = torch.Generator().manual_seed(1)
    for s in range(3):
        for p, q in zip(a.parameters(), b.parameters()):
            gr = torch.randn(p.shape, generator=gen)
            p.grad.copy_(gr)
            q.grad.copy_(gr)
        cpu_kernels["radam"].clear()
        oa.clip_grad_norm_(0.5); oa.step()
        assert len(cpu_kernels["radam"]) == 1
        cpu_kernels["radam"].clear()
        ob.clip_grad_norm_(0.5); ob.step()
        assert len(cpu_kernels["radam"]) == 4            # flows[1], flows[0], embedding, encoder (rest bucket: 2 runs)
        assert ob._flats[0].g is None
        for p, q in zip(a.parameters(), b.parameters()):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-6, atol=1e-8)
