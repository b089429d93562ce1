"""Time the UNMODIFIED reference (oracle/ref_shims.py: /root/reference here, its byte-compiled staging oracle/_ref/ on
the GPU box) on host CPU cores, fp32 — the `--impl reference` arm and the `cpu_baseline` leg of bench.py
(SURVEY.md §8d "Reference CPU path timed beside it").  Test infrastructure: never imported by
flowtron_b200/.

Workloads mirror bench.py's GPU arms on a BOUNDED sample:
  train : 2-flow Flowtron.forward + FlowtronLoss + backward (train.py:294-305; no optimizer, like BASELINE.md §2) at the
          cfg-2 sequence shape (T=1000, LJS-like text lengths, beta-binomial prior) with a batch small enough to finish
          in seconds (the [B,T,L,640] score tensor is 12 GB at B=32);
  infer : Flowtron.infer (flowtron.py:901-930), B=1 (the reference raises for B>1 with a gate layer), sigma=0.5;
  mel   : TacotronSTFT.mel_spectrogram one utterance at a time, as data.py:149-155 does.
Thread count: torch.set_num_threads(n) for n in a small candidate set, one short calibration step each, fastest wins
(128-way oversubscription of a 2-utterance batch was the round-1 bug: 1.76 frames/s).
"""
from __future__ import annotations

import os
import time

import torch

from . import ref_shims
from flowtron_b200 import synth


def physical_cores() -> int:
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            return int(n)
    except Exception:
        pass
    return max(1, (os.cpu_count() or 2) // 2)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def thread_candidates():
    phys, logical = physical_cores(), os.cpu_count() or 1
    return sorted({n for n in (8, 16, 32, phys) if 1 <= n <= logical})


def _pick_threads(step, cands):
    best, best_t, table = None, None, {}
    for n in cands:
        torch.set_num_threads(n)
        step()                                        # touch the allocator / thread pool at this width
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        table[n] = dt
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best, table


def _train_problem(B, T, seed=1234):
    cfg = dict(synth.DEFAULT_MODEL_CONFIG)
    F, model = ref_shims.reference_model(cfg, synth.synth_params(cfg, seed))
    model.train()                                     # like train.py (encoder dropout active)
    out_lens, in_lens = synth.ljs_like_lengths(B, T, seed)
    L = int(in_lens.max())
    batch = synth.synth_batch(B, T, L, cfg, seed, out_lens=out_lens.tolist(), in_lens=in_lens.tolist(), with_prior=True,
                              logmel_stats=True)
    crit = F.FlowtronLoss(sigma=1.0, gm_loss=False, gate_loss=True, use_ctc_loss=False)

    def step():
        for p in model.parameters():
            p.grad = None
        out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"], batch["attn_prior"])
        nll, gl, _ = crit(out, batch["gate_target"], batch["in_lens"], batch["out_lens"])
        (nll + gl).sum().backward()
    return step, int(batch["out_lens"].sum()), L


def time_train(steps=3, warmup=1, B=2, T=1000, threads=None, calibrate_T=250):
    """Returns dict(value = valid mel-frames/s, ...).  One step = forward + loss + backward of the reference."""
    if threads is None:
        cal_step, _, _ = _train_problem(B, calibrate_T)
        threads, table = _pick_threads(cal_step, thread_candidates())
    else:
        table = {}
        torch.set_num_threads(threads)
    step, frames, L = _train_problem(B, T)
    for _ in range(max(0, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": frames * steps / dt, "unit": "valid mel-frames/s", "cores": threads, "kind": "reference",
            "s_per_step": dt / steps, "thread_sweep_s": {str(k): round(v, 3) for k, v in table.items()},
            "cpu_model": cpu_model(), "physical_cores": physical_cores(),
            "sample": f"unmodified reference Flowtron (2 flows) fwd+loss+bwd, fp32, B={B}, T<={T}, L={L}, prior on, {steps} steps "
                      f"after {warmup} warm-up, torch threads={threads}"}


def time_infer(steps=3, warmup=1, T=400, L=100, threads=None):
    cfg = dict(synth.DEFAULT_MODEL_CONFIG)
    params = synth.synth_params(cfg, 1234)
    key = [k for k in params if k.endswith("gate_layer.linear_layer.bias")][0]
    params[key] = torch.full_like(params[key], -10.0)          # the gate never fires: fixed work (SURVEY §8d cfg 4)
    F, model = ref_shims.reference_model(cfg, params)
    g = torch.Generator().manual_seed(1234)
    residual = torch.randn(1, 80, T, generator=g) * 0.5
    text = torch.randint(0, cfg["n_text"], (1, L), generator=g)
    spk = torch.zeros(1, dtype=torch.long)

    def step(Tn=T):
        with torch.no_grad():
            model.infer(residual[:, :, :Tn], spk, text, temperature=1.0, gate_threshold=0.5)
    table = {}
    if threads is None:
        threads, table = _pick_threads(lambda: step(32), sorted(set([1, 4] + thread_candidates())))
    torch.set_num_threads(threads)
    for _ in range(max(0, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    fps = T * steps / dt
    return {"value": fps, "unit": "mel-frames/s", "cores": threads, "kind": "reference", "s_per_step": dt / steps,
            "rtf": fps / (22050.0 / 256.0), "thread_sweep_s": {str(k): round(v, 3) for k, v in table.items()},
            "cpu_model": cpu_model(), "physical_cores": physical_cores(),
            "sample": f"unmodified reference Flowtron.infer, 2 flows, B=1, T={T}, L={L}, sigma=0.5, fp32, {steps} requests, "
                      f"torch threads={threads}"}


def time_mel(n_utt=64, seed=7, threads=None):
    """Per-utterance TacotronSTFT.mel_spectrogram like the reference's data loader (data.py:149-155)."""
    AP = ref_shims.import_audio_processing()
    stft = AP.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0)
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(22050, 220501, (n_utt,), generator=g).tolist()
    wavs = [(torch.rand(1, n, generator=g) * 1.9 - 0.95) for n in lens]

    def run(ws):
        f = 0
        for w in ws:
            f += stft.mel_spectrogram(w).shape[-1]
        return f
    table = {}
    if threads is None:
        threads, table = _pick_threads(lambda: run(wavs[:4]), sorted(set([1, 4] + thread_candidates())))
    torch.set_num_threads(threads)
    run(wavs[:2])
    t0 = time.perf_counter()
    frames = run(wavs)
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": "mel-frames/s", "cores": threads, "kind": "reference",
            "gbs": frames * 1344.0 / dt / 1e9, "thread_sweep_s": {str(k): round(v, 3) for k, v in table.items()},
            "cpu_model": cpu_model(), "physical_cores": physical_cores(),
            "sample": f"unmodified reference TacotronSTFT.mel_spectrogram, {n_utt} utterances U{{1..10}} s one at a time "
                      f"(data.py:149-155), {frames} frames, fp32, torch threads={threads}"}
