"""GPU: persistent tcgen05 LSTM recurrence (forward + BPTT) vs the CPU oracle, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu

H = 1024


def _mk(T, B, seed, wscale=1.0):
    g = torch.Generator().manual_seed(seed)
    xproj = torch.randn(T, B, 4 * H, generator=g) * 0.8
    whh = torch.randn(4 * H, H, generator=g) * (wscale / (3 * H) ** 0.5)
    return xproj, whh


def _oracle(xproj, whh, lens):
    from oracle import flowtron_oracle as O
    T, B, _ = xproj.shape
    eye = torch.zeros(4 * H, 1)
    # lstm_layer_explicit computes x @ w_ih.T + b; feed xproj through a zero-width trick: use bias-free identity
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    outs, cs = [], []
    for t in range(T):
        a = xproj[t] + h @ whh.t()
        i, f, gg, o = a.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
        cs.append(c)
    out = torch.stack(outs)
    if lens is not None:
        m = (torch.arange(T)[:, None] < lens[None, :]).float()[:, :, None]
        out = out * m
    return out, torch.stack(cs)


@pytest.mark.parametrize("T,B,lens", [(6, 2, None), (40, 5, [40, 3, 17, 40, 1]), (33, 32, "rand"), (20, 48, "rand")])
def test_lstm_fwd_bwd(T, B, lens):
    from flowtron_b200 import _lib
    xproj, whh = _mk(T, B, T * 100 + B, wscale=2.0)
    if lens == "rand":
        g = torch.Generator().manual_seed(5)
        lens_t = torch.randint(1, T + 1, (B,), generator=g)
        lens_t[0] = T
    elif lens is None:
        lens_t = None
    else:
        lens_t = torch.tensor(lens)
    xp = xproj.clone().requires_grad_(True)
    ref_h, ref_c = _oracle(xp, whh.half().float(), lens_t)     # same fp16-rounded weights: isolates kernel error

    dev = "cuda"
    hseq = torch.zeros(T, B, H + 64, device=dev, dtype=torch.float16)[:, :, :H]   # strided view like d[T,B,1664]
    gates = torch.zeros(T, B, 4 * H, device=dev, dtype=torch.float16)
    cst = torch.zeros(T, B, H, device=dev)
    lens_d = None if lens_t is None else lens_t.to(dev, torch.int32)
    _lib.lstm_fwd(xproj.to(dev), whh.to(dev).half(), lens_d, hseq, gates, cst)
    torch.cuda.synchronize()
    assert _lib.device_status() == 0
    err = (hseq.float().cpu() - ref_h.detach()).abs().max().item()
    assert err < 3e-3, err
    vm = torch.ones(T, B, dtype=torch.bool) if lens_t is None else (torch.arange(T)[:, None] < lens_t[None, :])
    cerr = (cst.cpu() - ref_c.detach())[vm].abs().max().item()
    assert cerr < 5e-3, cerr

    # BPTT: d(sum(h * w))/d(xproj) == dG
    g = torch.Generator().manual_seed(9)
    w = torch.randn(T, B, H, generator=g)
    (ref_h * w).sum().backward()
    ref_dG = xp.grad
    dG = torch.zeros(T, B, 4 * H, device=dev, dtype=torch.float16)
    _lib.lstm_bwd(w.to(dev), whh.t().contiguous().to(dev).half(), gates, cst, lens_d, dG)
    torch.cuda.synchronize()
    assert _lib.device_status() == 0
    d = dG.float().cpu()
    scale = ref_dG.abs().max().item()
    rel = (d - ref_dG).norm().item() / ref_dG.norm().item()
    assert rel < 3e-3, rel
    assert (d - ref_dG).abs().max().item() < 1e-2 * scale
    if lens_t is not None:
        assert d[~vm].abs().max().item() == 0.0
