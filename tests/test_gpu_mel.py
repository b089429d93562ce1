"""GPU: TacotronSTFT mel front-end (cuFFT + fused mag/filterbank/log) vs fixtures generated from the reference's
own TacotronSTFT (conv-DFT route).  Tolerance 1e-3 of the log-mel range per north_star; measured ~1e-5."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

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
        assert err <= 1e-3 * max(1.0, float(ref.abs().max())), (k, err)
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
        assert torch.allclose(a, b, atol=1e-6)
    assert _lib.device_status() == 0
