"""Oracle for the long-form and streaming modes.  TEST INFRASTRUCTURE ONLY.

hop == window restates the reference's behaviour: src/datasets/ami/utils.py:107,163 (5 s windows, <= 3 s dropped,
padded to 5 s), src/engines/vad_engine.py:204-211 (per-window forward + median filter), src/scripts/predict.py:434,
447-458, 472-490 (flatten, slice ceil(duration / fs) + 1 frames, RLE).  hop < window and streaming have no reference
counterpart; their definitions are restated here in numpy so the CUDA path can be checked
bit-for-bit on the integer / index work.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from .postproc import median_filter, rle_segments, slice_recordings


def cut_windows(num_samples: int, window: int = 80000, hop: int = 80000):
    """[(start, valid)] -- reference windows for hop == window, plain sliding windows otherwise."""
    out = []
    if hop == window:
        for s in range(0, num_samples, window):
            n = min(window, num_samples - s)
            if n > 3 * 16000 * window // 80000:            # .filter(duration > 3) on 5 s windows
                out.append((s, n))
        return out
    s = 0
    while True:
        out.append((s, min(window, num_samples - s)))
        if s + window >= num_samples:
            break
        s += hop
    return out


def stitch_center(prob: np.ndarray, hop_frames: int, total_frames: int) -> np.ndarray:
    """prob (W, Tw): frame f comes from window clamp((f - (Tw - hop) // 2) // hop, 0, W - 1); 0 where uncovered."""
    W, Tw = prob.shape
    margin = (Tw - hop_frames) // 2
    out = np.zeros(total_frames, dtype=prob.dtype)
    for f in range(total_frames):
        w = 0 if f < margin else min((f - margin) // hop_frames, W - 1)
        k = f - w * hop_frames
        if 0 <= k < Tw:
            out[f] = prob[w, k]
    return out


def longform_reference(model, fbank_fn, wav: torch.Tensor, window: int = 80000, frame_shift: float = 0.01):
    """hop == window: the reference pipeline on one recording.  Returns (window probabilities, intervals)."""
    N = wav.numel()
    wins = cut_windows(N, window, window)
    rows = torch.zeros(len(wins), window)
    feats = []
    for i, (s, n) in enumerate(wins):
        f = fbank_fn(wav[s:s + n].unsqueeze(0))[0]                      # features of the cut itself
        Tw = (window + 80) // 160
        if f.shape[0] < Tw:                                             # .pad(duration=5.0): LOG_EPSILON feature padding
            f = torch.cat([f, torch.full((Tw - f.shape[0], f.shape[1]), -23.025850929940457)])
        feats.append(f)
    feats = torch.stack(feats)
    with torch.no_grad():
        prob = model(feats).squeeze(-1)
    dec = median_filter(prob, window=frame_shift)
    flat = dec.reshape(-1)
    stream = slice_recordings(flat, [N / 16000.0], frame_shift)[0]
    return prob, rle_segments(stream.tolist(), frame_shift)
