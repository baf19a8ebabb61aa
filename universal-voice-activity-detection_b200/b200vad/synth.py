"""Synthetic 16 kHz audio of the BASELINE.json shapes (there is no network for corpora).

"meeting-style": alternating speech-like bursts (3-5 harmonics of f0 in [90, 250] Hz under a slow
amplitude modulation, amplitude ~0.1) and background noise (1e-3..1e-2); burst / gap lengths
~Exp(mean 2 s / 1 s).  ``noise_batch`` is plain 0.1*randn for pure-throughput runs.
"""

from __future__ import annotations

import math

import torch


def noise_batch(B: int, N: int, seed: int = 1234, device="cpu", pin: bool = False) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty((B, N), dtype=torch.float32, device=device, pin_memory=pin and str(device) == "cpu")
    x.normal_(0.0, 0.1, generator=g)
    return x


def meeting_batch(B: int, N: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    """(B, N) float32 in roughly [-0.5, 0.5]; deterministic per (seed, row)."""
    out = torch.empty((B, N), dtype=torch.float32)
    t = torch.arange(N, dtype=torch.float64) / 16000.0
    for b in range(B):
        g = torch.Generator().manual_seed(seed + b)
        noise_amp = 10 ** (-3 + torch.rand(1, generator=g).item())
        x = noise_amp * torch.randn(N, generator=g, dtype=torch.float64)
        pos = 0.0
        speech = torch.rand(1, generator=g).item() < 0.5
        dur = N / 16000.0
        while pos < dur:
            mean = 2.0 if speech else 1.0
            length = -mean * math.log(max(torch.rand(1, generator=g).item(), 1e-6))
            a, bnd = int(pos * 16000), min(N, int((pos + length) * 16000))
            if speech and bnd > a:
                f0 = 90 + 160 * torch.rand(1, generator=g).item()
                nh = 3 + int(torch.randint(0, 3, (1,), generator=g).item())
                seg_t = t[a:bnd]
                am = 0.5 * (1 + torch.sin(2 * math.pi * (2 + 3 * torch.rand(1, generator=g).item()) * seg_t))
                s = torch.zeros_like(seg_t)
                for h in range(1, nh + 1):
                    s += torch.sin(2 * math.pi * f0 * h * seg_t + 6.28 * torch.rand(1, generator=g).item()) / h
                x[a:bnd] += 0.1 * am * s
            pos += length
            speech = not speech
        out[b] = x.float()
    return out.to(device)
