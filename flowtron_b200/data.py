"""Device side of the reference's data layer for the hot path (data.py).

The reference computes the beta-binomial attention prior with one ``scipy.stats.betabinom.pmf`` call per mel frame on a
CPU DataLoader worker (data.py:31-41), zero-pads everything in ``DataCollate`` (data.py:191-246) and ships the padded
tensors to the GPU every step — the [B,T,L] prior alone is 19.7 MB of the 30.1 MB cfg-2 step input.  Here the prior and
the padding are produced on the device, from lengths alone / from the packed output of the GPU mel front-end:

    beta_binomial_prior_distribution(P, M, s)   same name and result as data.py:31-41, on the GPU
    attn_prior_batch(in_lens, out_lens, T, L)   the whole padded [B,T,L] tensor in one kernel (what bench.py's e2e uses)
    DataCollate(...)                            data.py:191-246 for GPU-resident items (mel [80,F] CUDA tensors or a packed
                                                mel + frame offsets), returning the same 7-tuple on the device

Text cleaning / ARPAbet / wav file IO stay outside the scope (out of the hot path).  CUDA only: CPU tensors
raise (no fallback).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import FlowtronB200Error


def attn_prior_batch(in_lens: torch.Tensor, out_lens: torch.Tensor, T: int, L: int, scaling_factor: float = 1.0,
                     threshold: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,T,L] zero-padded beta-binomial prior (data.py:31-41, threshold :137-139, padding :238-243) from lengths."""
    if not in_lens.is_cuda:
        raise FlowtronB200Error("attn_prior_batch needs CUDA tensors")
    dev = in_lens.device
    il = in_lens.to(torch.int32).contiguous()
    ol = out_lens.to(device=dev, dtype=torch.int32).contiguous()
    B = il.numel()
    if out is None:
        out = torch.empty(B, T, L, device=dev)
    _lib.attn_prior(il, ol, T, L, scaling_factor, threshold, out)
    return out


def beta_binomial_prior_distribution(phoneme_count: int, mel_count: int, scaling_factor: float = 1.0, device="cuda"):
    """data.py:31-41: [mel_count, phoneme_count] (the reference returns float64 from scipy; this is its fp32 cast, which is
    what DataCollate stores, data.py:222)."""
    dev = torch.device(device)
    il = torch.tensor([phoneme_count], dtype=torch.int32, device=dev)
    ol = torch.tensor([mel_count], dtype=torch.int32, device=dev)
    return attn_prior_batch(il, ol, mel_count, phoneme_count, scaling_factor)[0]


class DataCollate:
    """data.py:191-246 on the device.  ``batch`` = list of (mel [n_mel, F] CUDA, speaker_id, text_encoded 1-D long,
    attn_prior or None).  Returns (mel_padded, speaker_ids, text_padded, input_lengths, output_lengths, gate_padded,
    attn_prior_padded) with the reference's shapes, dtypes and row order (sorted by text length, descending, stable).
    With ``use_attn_prior`` the prior is computed on the device from the lengths (items' own priors are ignored)."""

    def __init__(self, n_frames_per_step=1, use_attn_prior=False, betab_scaling_factor=1.0, attn_prior_threshold=0.0):
        self.n_frames_per_step = n_frames_per_step
        self.use_attn_prior = use_attn_prior
        self.betab_scaling_factor = betab_scaling_factor
        self.attn_prior_threshold = attn_prior_threshold

    def __call__(self, batch):
        mels = [b[0] for b in batch]
        if not mels[0].is_cuda:
            raise FlowtronB200Error("DataCollate (device) needs CUDA mel tensors")
        n_mel = mels[0].size(0)
        frames = [int(m.size(1)) for m in mels]
        packed = torch.cat([m.reshape(-1).float() for m in mels])
        fo = torch.tensor([0] + list(torch.tensor(frames).cumsum(0).tolist()), dtype=torch.int64)
        return self.collate_packed(packed, fo, n_mel, [b[1] for b in batch], [b[2] for b in batch])

    def collate_packed(self, mel_packed: torch.Tensor, frame_offsets: torch.Tensor, n_mel: int, speaker_ids: Sequence,
                       texts: Sequence[torch.Tensor]):
        """Same, from the packed output of TacotronSTFT.mel_spectrogram_packed (no per-utterance tensors at all)."""
        dev = mel_packed.device
        B = len(texts)
        fo_host = torch.as_tensor(frame_offsets, dtype=torch.int64).cpu()
        frames = (fo_host[1:] - fo_host[:-1])
        text_lens = torch.tensor([int(t.numel()) for t in texts], dtype=torch.long)
        input_lengths, order = torch.sort(text_lens, dim=0, descending=True, stable=True)     # data.py:200-202
        max_input_len = int(input_lengths[0])
        text_padded = torch.zeros(B, max_input_len, dtype=torch.long)
        for i, j in enumerate(order.tolist()):
            text_padded[i, :texts[j].numel()] = texts[j].reshape(-1).cpu()
        max_target_len = int(frames.max())
        if max_target_len % self.n_frames_per_step != 0:
            max_target_len += self.n_frames_per_step - max_target_len % self.n_frames_per_step
        mel_padded = torch.empty(B, n_mel, max_target_len, device=dev)
        gate_padded = torch.empty(B, max_target_len, device=dev)
        out_lens32 = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.collate_mel(mel_packed, fo_host.to(dev), order.to(device=dev, dtype=torch.int32), n_mel, max_target_len, mel_padded,
                         gate_padded, out_lens32)
        spk = torch.tensor([int(speaker_ids[j]) for j in order.tolist()], dtype=torch.long, device=dev)
        input_lengths_d = input_lengths.to(dev)
        prior = None
        if self.use_attn_prior:
            prior = attn_prior_batch(input_lengths_d, out_lens32, max_target_len, max_input_len, self.betab_scaling_factor,
                                     self.attn_prior_threshold)
        return (mel_padded, spk, text_padded.to(dev), input_lengths_d, out_lens32.long(), gate_padded, prior)
