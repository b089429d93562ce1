"""Deterministic synthetic weights and inputs (the reference ships no data: filelists point at
/path_to_ljs/...).  Shared by bench.py, __graft_entry__.smoke(), the golden
generator and the tests.

Weights are a pure function of (parameter name, shape, seed) — independent of module construction
order — so the reference model (in tests/make_golden.py), the oracle and the CUDA path all see
identical values without shipping a 240 MB checkpoint.  The key layout is the reference's
state_dict layout (train.py:131-139).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import numpy as np
import torch

DEFAULT_MODEL_CONFIG = dict(  # /root/reference/config.json:49-66
    n_speakers=1, n_speaker_dim=128, n_text=185, n_text_dim=512, n_flows=2, n_mel_channels=80,
    n_attn_channels=640, n_hidden=1024, n_lstm_layers=2, mel_encoder_n_hidden=512, n_components=0,
    mean_scale=0.0, fixed_gaussian=True, dummy_speaker_embedding=False, use_gate_layer=True,
    use_cumm_attention=False)


def param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape for Flowtron(**cfg) (n_components=0, no cumulative attention)."""
    H, A, M = cfg["n_hidden"], cfg["n_attn_channels"], cfg["n_mel_channels"]
    Dt, Ds = cfg["n_text_dim"], cfg["n_speaker_dim"]
    E = Dt + Ds
    s: Dict[str, Tuple[int, ...]] = {}
    s["speaker_embedding.weight"] = (cfg["n_speakers"], Ds)
    s["embedding.weight"] = (cfg["n_text"], Dt)
    for i in range(3):
        s[f"encoder.convolutions.{i}.0.conv.weight"] = (Dt, Dt, 5)
        s[f"encoder.convolutions.{i}.0.conv.bias"] = (Dt,)
        s[f"encoder.convolutions.{i}.1.weight"] = (Dt,)
        s[f"encoder.convolutions.{i}.1.bias"] = (Dt,)
    for sfx in ("", "_reverse"):
        s[f"encoder.lstm.weight_ih_l0{sfx}"] = (4 * (Dt // 2), Dt)
        s[f"encoder.lstm.weight_hh_l0{sfx}"] = (4 * (Dt // 2), Dt // 2)
        s[f"encoder.lstm.bias_ih_l0{sfx}"] = (4 * (Dt // 2),)
        s[f"encoder.lstm.bias_hh_l0{sfx}"] = (4 * (Dt // 2),)
    for i in range(cfg["n_flows"]):
        pre = f"flows.{i}" if i % 2 == 0 else f"flows.{i}.ar_step"
        s[f"{pre}.conv.weight"] = (2 * M, H, 1)
        s[f"{pre}.conv.bias"] = (2 * M,)
        for l in range(cfg["n_lstm_layers"]):
            s[f"{pre}.lstm.weight_ih_l{l}"] = (4 * H, (H + A) if l == 0 else H)
            s[f"{pre}.lstm.weight_hh_l{l}"] = (4 * H, H)
            s[f"{pre}.lstm.bias_ih_l{l}"] = (4 * H,)
            s[f"{pre}.lstm.bias_hh_l{l}"] = (4 * H,)
        s[f"{pre}.attention_lstm.weight_ih_l0"] = (4 * H, M)
        s[f"{pre}.attention_lstm.weight_hh_l0"] = (4 * H, H)
        s[f"{pre}.attention_lstm.bias_ih_l0"] = (4 * H,)
        s[f"{pre}.attention_lstm.bias_hh_l0"] = (4 * H,)
        s[f"{pre}.attention_layer.query.linear_layer.weight"] = (A, H)
        s[f"{pre}.attention_layer.key.linear_layer.weight"] = (A, E)
        s[f"{pre}.attention_layer.value.linear_layer.weight"] = (A, E)
        s[f"{pre}.attention_layer.v.linear_layer.weight"] = (1, A)
        for j in range(2):
            s[f"{pre}.dense_layer.layers.{j}.linear_layer.weight"] = (H, H)
            s[f"{pre}.dense_layer.layers.{j}.linear_layer.bias"] = (H,)
        if i == cfg["n_flows"] - 1 and cfg["use_gate_layer"]:
            s[f"{pre}.gate_layer.linear_layer.weight"] = (1, H + A)
            s[f"{pre}.gate_layer.linear_layer.bias"] = (1,)
    return s


def _std_for(name: str, shape) -> float:
    if name.endswith("embedding.weight"):
        return 1.0
    if ".conv.weight" in name and name.startswith("flows"):
        return 0.02                       # non-trivial affine coupling (reference zero-inits, flowtron.py:652)
    if ".conv.bias" in name and name.startswith("flows"):
        return 0.02
    if "lstm" in name:                    # torch LSTM default U(-k,k), k=1/sqrt(H): std = k/sqrt(3)
        H = shape[0] // 4
        return 1.0 / math.sqrt(3.0 * H)
    if name.endswith(".1.weight"):        # instance-norm affine weight (offset +1 applied below)
        return 0.1
    if name.endswith("bias"):
        return 0.05
    if len(shape) >= 2:                   # xavier-like
        fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
        return math.sqrt(2.0 / (fan_in + fan_out))
    return 0.05


def synth_params(cfg: dict, seed: int = 1234, scale: float = 1.0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic N(0, std(name)) weights for every state_dict key."""
    out = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator(device="cpu")
        g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
        t = torch.randn(shape, generator=g, dtype=torch.float32) * (_std_for(name, shape) * scale)
        if name.endswith(".1.weight"):
            t = t + 1.0
        out[name] = t.to(dtype)
    return out


def synth_batch(B: int, T: int, L: int, cfg: dict = DEFAULT_MODEL_CONFIG, seed: int = 1234,
                out_lens=None, in_lens=None, with_prior: bool = False, logmel_stats: bool = False):
    """Synthetic training batch in DataCollate layout (data.py:197-246):
    mel [B,80,T] f32, speaker_ids [B], text [B,L] i64, in_lens [B] (sorted desc), out_lens [B],
    gate_target [B,T], attn_prior [B,T,L] | None.  Pads are zero like the collate's."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    if out_lens is None:
        lo = max(1, T // 2)
        out_lens = torch.randint(lo, T + 1, (B,), generator=g)
        out_lens[int(torch.randint(0, B, (1,), generator=g))] = T
    else:
        out_lens = torch.as_tensor(out_lens, dtype=torch.long)
    if in_lens is None:
        in_lens = torch.clamp((out_lens.float() / T * L).round().long(), min=max(1, L // 4), max=L)
        in_lens[out_lens.argmax()] = L
    else:
        in_lens = torch.as_tensor(in_lens, dtype=torch.long)
    order = torch.argsort(in_lens, descending=True, stable=True)      # collate sorts by text length
    in_lens, out_lens = in_lens[order].contiguous(), out_lens[order].contiguous()
    M = cfg["n_mel_channels"]
    mel = torch.randn(B, M, T, generator=g)
    if logmel_stats:
        mel = (mel * 2.0 - 5.5).clamp_(-11.5, 2.0)
    text = torch.randint(0, cfg["n_text"], (B, L), generator=g)
    spk = torch.randint(0, cfg["n_speakers"], (B,), generator=g)
    tmask = torch.arange(T)[None, :] < out_lens[:, None]
    lmask = torch.arange(L)[None, :] < in_lens[:, None]
    mel = mel * tmask[:, None, :]
    text = text * lmask
    gate = (torch.arange(T)[None, :] >= (out_lens[:, None] - 1)).float()
    prior = None
    if with_prior:
        prior = torch.zeros(B, T, L)
        for b in range(B):
            pb = beta_binomial_prior(int(in_lens[b]), int(out_lens[b]))
            prior[b, : int(out_lens[b]), : int(in_lens[b])] = torch.from_numpy(pb).float()
    return dict(mel=mel, speaker_ids=spk, text=text, in_lens=in_lens, out_lens=out_lens,
                gate_target=gate, attn_prior=prior)


def beta_binomial_prior(phoneme_count: int, mel_count: int, scaling: float = 1.0) -> np.ndarray:
    """Attention prior of data.py:31-41 (BetaBinomial(k; n=P-1, a=s*i, b=s*(M+1-i)), i=1..M) in closed form via
    log-gamma; float64 [mel_count, phoneme_count].  (The reference loops scipy.stats.betabinom per frame.)"""
    from scipy.special import gammaln
    P, Mx = phoneme_count, mel_count
    k = np.arange(P, dtype=np.float64)[None, :]
    i = np.arange(1, Mx + 1, dtype=np.float64)[:, None]
    a, b, n = scaling * i, scaling * (Mx + 1 - i), float(P - 1)
    logc = gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1)
    logb = (gammaln(k + a) + gammaln(n - k + b) - gammaln(n + a + b)
            - (gammaln(a) + gammaln(b) - gammaln(a + b)))
    return np.exp(logc + logb)


def ljs_like_lengths(B: int, T: int, seed: int, lo_frac: float = 0.5):
    """cfg-2 style lengths: out_lens ~ U{lo..T} with the max forced to T,
    in_lens = round(out_lens / 6.5) clipped to [40, 160]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out_lens = torch.randint(int(T * lo_frac), T + 1, (B,), generator=g)
    out_lens[0] = T
    in_lens = torch.clamp((out_lens.float() / 6.5).round().long(), 40, 160)
    return out_lens, in_lens
