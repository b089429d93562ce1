#!/usr/bin/env python
"""Secondary measurements (BASELINE.json configs[3] and [4]): autoregressive inference frames/s + RTF, and the mel
front-end sweep (frames/s, GB/s of algorithmic bytes).  Prints one JSON line each; numbers go to profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowtron_b200 import _lib, synth
from flowtron_b200.flowtron import Flowtron
from flowtron_b200.audio_processing import TacotronSTFT


def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def infer():
    cfg = dict(synth.DEFAULT_MODEL_CONFIG)
    p = synth.synth_params(cfg, 1234)
    key = [k for k in p if k.endswith("gate_layer.linear_layer.bias")][0]
    p[key] = torch.full_like(p[key], -10.0)                       # never stop: fixed work (SURVEY §8d cfg 4)
    m = Flowtron(**cfg); m.load_state_dict(p, strict=True); m = m.cuda().eval()
    g = torch.Generator().manual_seed(1)
    for B, T in ((1, 400), (16, 400), (1, 1000)):
        L = 100
        z = (torch.randn(B, 80, T, generator=g) * 0.5).cuda()
        text = torch.randint(0, 185, (B, L), generator=g).cuda()
        spk = torch.zeros(B, dtype=torch.long, device="cuda")
        with torch.no_grad():
            ms = ev_time(lambda: m.infer(z, spk, text))
        fps = B * T / (ms / 1e3)
        print(json.dumps({"metric": "inference mel-frames/sec", "config": f"2-flow Flowtron.infer, sigma=0.5, B={B}, T={T}, L={L}",
                          "value": fps, "ms": ms, "rtf_x_realtime": fps / 86.13, "us_per_frame_per_flow": ms * 1e3 / T / 2}))


def mel():
    stft = TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).cuda()
    g = torch.Generator().manual_seed(2)
    n_utt = 2000                                                  # 1/5 of cfg 5's 10k utterances (same length law)
    lens = (torch.randint(1, 11, (n_utt,), generator=g) * 22050).tolist()
    wavs = [(torch.rand(n, generator=g) * 1.9 - 0.95).cuda() for n in lens]
    frames = sum(1 + n // 256 for n in lens)
    ms = ev_time(lambda: stft.mel_spectrogram_ragged(wavs), n=3)
    fps = frames / (ms / 1e3)
    print(json.dumps({"metric": "mel front-end frames/sec", "config": f"{n_utt} utterances U{{1..10}}s @22.05kHz, hop 256, 80 mel (incl. torch.cat of the ragged list)",
                      "value": fps, "ms": ms, "frames": frames, "algorithmic_GBps": fps * 1344 / 1e9,
                      "hbm_frac_of_measured_peak": fps * 1344 / 1e9 / 6561.6}))


if __name__ == "__main__":
    infer()
    mel()
