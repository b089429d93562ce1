"""GPU: TacotronSTFT mel front-end vs fixtures generated from the reference's own TacotronSTFT (conv-DFT route): the fused
single-kernel path (n_fft 1024: per-warp shared-memory FFT, csrc/mel_fused.cu) through the module API and the general
cuFFT path (csrc/mel.cu) through the C ABI.  Tolerance 1e-3 of the log-mel range per north_star; measured ~1e-5."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, record_parity

pytestmark = pytest.mark.gpu


def _stft():
    from flowtron_b200.audio_processing import TacotronSTFT
    return TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).cuda()


def test_mel_matches_reference_goldens():
    from flowtron_b200 import _lib
    gold = dict(np.load(os.path.join(GOLDEN, "mel.npz")))
    stft = _stft()
    for k in ("demo", "noise", "short", "quiet"):
        y = torch.from_numpy(gold[f"y_{k}"]).cuda()
        m = stft.mel_spectrogram(y)
        torch.cuda.synchronize()
        ref = torch.from_numpy(gold[f"mel_{k}"])
        assert tuple(m.shape) == tuple(ref.shape)
        err = (m.cpu() - ref).abs().max().item()
        print(k, "abs err", err, "range", float(ref.max() - ref.min()))
        record_parity(f"mel_fused_{k}", {"abs_err": err, "ref_absmax": float(ref.abs().max())})
        assert err <= 1e-3 * max(1.0, float(ref.abs().max())), (k, err)
        # the general (cuFFT) path through the C ABI: same fixtures
        B, N = y.shape
        F = 1 + N // 256
        so = torch.arange(B + 1, device="cuda", dtype=torch.int64) * N
        fo = torch.arange(B + 1, device="cuda", dtype=torch.int64) * F
        m2 = torch.empty(B, 80, F, device="cuda")
        _lib.mel_spectrogram(y.contiguous(), so, fo, B, B * F, stft.stft_fn._window, stft.mel_basis, stft._band_lo, stft._band_hi,
                             1024, 256, 1e-5, m2)
        err2 = (m2.cpu() - ref).abs().max().item()
        assert err2 <= 1e-3 * max(1.0, float(ref.abs().max())), (k, err2)
    assert _lib.device_status() == 0


def test_ragged_equals_per_utterance_and_chunking():
    """Ragged batch (different lengths, chunk boundary inside an utterance) == one utterance at a time."""
    from flowtron_b200 import _lib
    stft = _stft()
    g = torch.Generator().manual_seed(0)
    lens = [22050, 9000, 1300, 40000, 600]
    wavs = [(torch.rand(n, generator=g) * 1.9 - 0.95).cuda() for n in lens]
    single = [stft.mel_spectrogram(w[None])[0] for w in wavs]
    ragged = stft.mel_spectrogram_ragged(wavs)
    for a, b in zip(single, ragged):
        assert a.shape == b.shape and torch.equal(a, b)
    # force tiny chunks through the C ABI: same answer
    import flowtron_b200._lib as L
    hop, n_fft = 256, 1024
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    fo = np.concatenate([[0], np.cumsum([1 + n // hop for n in lens])]).astype(np.int64)
    flat = torch.cat(wavs)
    out = torch.empty(int(fo[-1]) * 80, device="cuda")
    L.mel_spectrogram(flat, torch.from_numpy(so).cuda(), torch.from_numpy(fo).cuda(), len(lens), int(fo[-1]),
                      stft.stft_fn._window, stft.mel_basis, stft._band_lo, stft._band_hi, n_fft, hop, 1e-5, out, chunk_frames=37)
    for i, a in enumerate(single):
        b = out[fo[i] * 80: fo[i + 1] * 80].view(80, -1)
        assert torch.allclose(a, b, atol=2e-4)          # two different FFT algorithms (own radix-8 vs cuFFT), fp32
    assert _lib.device_status() == 0


def test_odd_offsets_int16_and_packed_api():
    """Odd utterance lengths put later utterances at odd sample offsets (scalar-load path of the fused kernel): results
    must equal the aligned single-utterance runs bit for bit.  int16 PCM input == the same samples given as f32/32768."""
    from flowtron_b200 import _lib
    stft = _stft()
    g = torch.Generator().manual_seed(1)
    lens = [22051, 9001, 1301, 40003, 777, 513]
    wavs = [(torch.rand(n, generator=g) * 1.9 - 0.95).cuda() for n in lens]
    single = [stft.mel_spectrogram(w[None])[0] for w in wavs]
    packed, fo = stft.mel_spectrogram_packed(torch.cat(wavs), lens)
    for i, a in enumerate(single):
        assert torch.equal(a, packed[fo[i] * 80: fo[i + 1] * 80].view(80, -1)), i
    pcm = [torch.round(w * 32767.0).to(torch.int16) for w in wavs]
    p16, _ = stft.mel_spectrogram_packed(torch.cat(pcm), lens)
    p32, _ = stft.mel_spectrogram_packed(torch.cat([p.float() / 32768.0 for p in pcm]), lens)
    assert (p16 - p32).abs().max().item() <= 1e-6
    assert _lib.device_status() == 0


def test_stft_transform_matches_reference_algorithm():
    """STFT.transform (audio_processing.py:207-235) vs its CPU restatement (pinned to the reference in
    tests/test_oracle_data.py): magnitude, and phase compared through the complex value it encodes (phase itself is
    ill-conditioned where the magnitude vanishes and wraps at +-pi)."""
    from flowtron_b200 import _lib
    from flowtron_b200.audio_processing import STFT
    from oracle import stft_oracle as S
    g = torch.Generator().manual_seed(3)
    y = torch.rand(3, 5000, generator=g) * 1.9 - 0.95
    m_ref, p_ref = S.stft_transform(y)
    m, p = STFT(1024, 256, 1024).cuda().transform(y.cuda())
    torch.cuda.synchronize()
    assert tuple(m.shape) == tuple(m_ref.shape) == (3, 513, 1 + 5000 // 256)
    scale = m_ref.abs().max().item()
    e_mag = (m.cpu() - m_ref).abs().max().item() / scale
    zr, zi = (m * torch.cos(p)).cpu(), (m * torch.sin(p)).cpu()
    e_cplx = max((zr - m_ref * torch.cos(p_ref)).abs().max().item(), (zi - m_ref * torch.sin(p_ref)).abs().max().item()) / scale
    record_parity("stft_transform", {"magnitude": e_mag, "complex": e_cplx})
    assert e_mag <= 1e-5 and e_cplx <= 1e-5, (e_mag, e_cplx)
    assert _lib.device_status() == 0
