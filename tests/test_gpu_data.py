"""GPU: the device-side data layer (flowtron_b200/data.py: ft_attn_prior, ft_collate_mel) vs the CPU restatement of the
reference's data.py (oracle/data_oracle.py, pinned to the reference in tests/test_oracle_data.py)."""
import numpy as np
import pytest
import torch

from conftest import record_parity
from oracle import data_oracle as D
from flowtron_b200 import synth

pytestmark = pytest.mark.gpu


def test_prior_matches_scipy_betabinom():
    from flowtron_b200 import _lib
    from flowtron_b200.data import attn_prior_batch, beta_binomial_prior_distribution
    cases = [(11, 29), (40, 200), (1, 7), (154, 1000), (79, 512)]
    in_lens = torch.tensor([c[0] for c in cases], dtype=torch.int32, device="cuda")
    out_lens = torch.tensor([c[1] for c in cases], dtype=torch.int32, device="cuda")
    T, L = 1000, 160
    pr = attn_prior_batch(in_lens, out_lens, T, L).cpu()
    torch.cuda.synchronize()
    worst = 0.0
    for b, (P, M) in enumerate(cases):
        ref = torch.from_numpy(synth.beta_binomial_prior(P, M)).float()          # == scipy loop (tests/test_oracle_data.py)
        got = pr[b, :M, :P]
        worst = max(worst, ((got - ref).abs() / (ref.abs() + 1e-30)).max().item() if ref.min() > 0 else (got - ref).abs().max().item())
        assert torch.allclose(got, ref, rtol=2e-6, atol=1e-37), (P, M)
        assert float(pr[b, M:].abs().sum()) == 0.0 and float(pr[b, :, P:].abs().sum()) == 0.0      # zero padding
    record_parity("attn_prior", {"worst_rel": worst})
    one = beta_binomial_prior_distribution(11, 29).cpu()
    assert torch.allclose(one, D.beta_binomial_prior_distribution(11, 29).float(), rtol=2e-6, atol=1e-37)
    thr = attn_prior_batch(in_lens, out_lens, T, L, threshold=1e-3).cpu()
    assert torch.equal(thr, torch.where(pr < 1e-3, torch.zeros_like(pr), pr))
    # scaling factor != 1 (non-integer beta parameters)
    half = attn_prior_batch(in_lens[:2], out_lens[:2], 200, 40, scaling_factor=0.5).cpu()
    assert torch.allclose(half[0, :29, :11], D.beta_binomial_prior_distribution(11, 29, 0.5).float(), rtol=5e-6, atol=1e-37)
    assert _lib.device_status() == 0


def test_device_collate_matches_reference_collate():
    from flowtron_b200 import _lib
    from flowtron_b200.data import DataCollate
    g = torch.Generator().manual_seed(0)
    shapes = [(29, 11), (40, 17), (5, 17), (33, 3), (40, 9)]
    cpu_batch, gpu_batch = [], []
    for i, (F_, P_) in enumerate(shapes):
        mel = torch.randn(80, F_, generator=g)
        text = torch.randint(1, 100, (P_,), generator=g)
        cpu_batch.append((mel, torch.LongTensor([i % 3]), text, D.beta_binomial_prior_distribution(P_, F_, 1.0)))
        gpu_batch.append((mel.cuda(), torch.LongTensor([i % 3]), text, None))
    ref = D.collate(cpu_batch, 1, use_attn_prior=True)
    got = DataCollate(1, use_attn_prior=True)(gpu_batch)
    torch.cuda.synchronize()
    names = ["mel_padded", "speaker_ids", "text_padded", "input_lengths", "output_lengths", "gate_padded", "attn_prior_padded"]
    for n, a, b in zip(names, ref, got):
        assert tuple(a.shape) == tuple(b.shape), n
        if n == "attn_prior_padded":
            assert torch.allclose(a, b.cpu(), rtol=2e-6, atol=1e-37), n
        else:
            assert torch.equal(a.float(), b.cpu().float()), n
    assert got[3].dtype == torch.long and got[4].dtype == torch.long and got[2].dtype == torch.long
    assert _lib.device_status() == 0


def test_wav_to_training_batch_on_device():
    """wav (packed, int16) -> fused mel -> device collate + prior: the batch the flows consume, with no host tensor but the
    lengths; checked against the per-piece CPU restatements."""
    from flowtron_b200.audio_processing import TacotronSTFT
    from flowtron_b200.data import DataCollate
    from oracle import stft_oracle as S
    g = torch.Generator().manual_seed(5)
    lens = [9000, 15000, 4000]
    wavs = [torch.rand(n, generator=g) * 1.9 - 0.95 for n in lens]
    texts = [torch.randint(1, 100, (p,), generator=g) for p in (12, 20, 7)]
    stft = TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).cuda()
    packed, fo = stft.mel_spectrogram_packed(torch.cat(wavs).cuda(), lens)
    out = DataCollate(1, use_attn_prior=True).collate_packed(packed, torch.from_numpy(fo), 80, [0, 0, 0], texts)
    torch.cuda.synchronize()
    mels = [S.mel_spectrogram(w[None])[0] for w in wavs]
    ref = D.collate([(m, 0, t, D.beta_binomial_prior_distribution(t.numel(), m.size(1))) for m, t in zip(mels, texts)], 1, True)
    assert torch.equal(ref[4], out[4].cpu()) and torch.equal(ref[3], out[3].cpu()) and torch.equal(ref[2], out[2].cpu())
    assert (ref[0] - out[0].cpu()).abs().max().item() <= 1e-3
    assert torch.equal(ref[5], out[5].cpu())
    assert torch.allclose(ref[6], out[6].cpu(), rtol=2e-6, atol=1e-37)
