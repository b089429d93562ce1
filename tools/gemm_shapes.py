#!/usr/bin/env python
"""Isolated timing of the training step's main GEMM shapes (CUDA events, L2 flushed between launches by rotating buffers),
single-CTA tiles vs CTA pairs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowtron_b200 import _lib

SHAPES = [  # name, M, N, K, a_mn, b_mn, out
    ("fwd xproj0", 32000, 4096, 1664, False, False, "f32"),
    ("fwd dense", 32000, 1024, 1024, False, False, "f16"),
    ("fwd xproj1 chunk", 3200, 4096, 1024, False, False, "f32"),
    ("dgrad ih0", 32000, 1664, 4096, False, True, "f32"),
    ("dgrad dense", 32000, 1024, 1024, False, True, "f16"),
    ("wgrad ih0", 4096, 1664, 32000, True, True, "f32"),
    ("wgrad hh", 4096, 1024, 32000, True, True, "f32"),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, M, N, K, a_mn, b_mn, out in SHAPES:
    if only and only not in name:
        continue
    nbuf = 3
    As = [(torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.3).half() for _ in range(nbuf)]
    Bs = [(torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.3).half() for _ in range(nbuf)]
    o32 = torch.empty(M, N, device="cuda") if out == "f32" else None
    o16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if out == "f16" else None
    res = []
    for mode in (0, 2):
        _lib.set_gemm_pair_mode(mode)
        for i in range(3):
            _lib.gemm(As[i % nbuf], Bs[i % nbuf], a_mn=a_mn, b_mn=b_mn, out32=o32, out16=o16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 12
        e0.record()
        for i in range(n):
            _lib.gemm(As[i % nbuf], Bs[i % nbuf], a_mn=a_mn, b_mn=b_mn, out32=o32, out16=o16)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res.append((ms, 2.0 * M * N * K / ms / 1e9))
    print(f"{name:18s} {M:6d} x {N:5d} x {K:6d}  single {res[0][0]:.3f} ms {res[0][1]:7.1f} TF/s | pair {res[1][0]:.3f} ms {res[1][1]:7.1f} TF/s")
_lib.set_gemm_pair_mode(1)
