"""Detection-error scoring on the GPU: the step right after segment extraction.

Reference: src/scripts/predict.py:500-509 (accumulation), :654-673 (get_binary_tensor, get_false_alarm,
get_missed_detection).  The host converts interval seconds to frame indices exactly as the reference does
(``int(t / frame_shift)`` in float64) and divides the GPU's integer frame counts the way the reference does
(torch integer sum / python int -> float32 tensor), so the printed rates are bit-identical.
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch


def intervals_to_frames(intervals_per_rec: Sequence[Sequence[Tuple[float, float]]], frame_shift: float) -> torch.Tensor:
    """[(start_s, end_s)] per recording -> (n, 3) int32 (recording, int(start/fs), int(end/fs)), predict.py:658."""
    rows = []
    for r, ivs in enumerate(intervals_per_rec):
        for start, end in ivs:
            rows.append((r, int(start / frame_shift), int(end / frame_shift)))
    if not rows:
        return torch.empty((0, 3), dtype=torch.int32)
    return torch.tensor(rows, dtype=torch.int32)


def detection_error(gt_intervals: Sequence[Sequence[Tuple[float, float]]], pred_intervals: Sequence[Sequence[Tuple[float, float]]],
                    durations: Sequence[float], frame_shift: float = 0.01, device="cuda"):
    """Per-recording and average (DER, FA, MD) as predict.py:500-509, 590-600 computes them.

    Returns dict(false_alarm, missed_detection, detection_error: float32 0-dim tensors averaged over recordings,
    fa_frames, md_frames: (R) int64 CPU tensors, nframes: list)."""
    R = len(durations)
    assert len(gt_intervals) == R and len(pred_intervals) == R
    nframes = [math.ceil(d / frame_shift) for d in durations]                 # predict.py:655
    words = [(n + 31) // 32 for n in nframes]
    word_off = [0]
    for w in words:
        word_off.append(word_off[-1] + w)
    dev = torch.device(device)
    fa, md = torch.ops.b200vad.score_intervals(
        intervals_to_frames(gt_intervals, frame_shift).to(dev), intervals_to_frames(pred_intervals, frame_shift).to(dev),
        torch.tensor(word_off, dtype=torch.int64, device=dev), torch.tensor(nframes, dtype=torch.int32, device=dev),
        word_off[-1], max(words) if words else 0)
    fa, md = fa.cpu(), md.cpu()
    fa_avg = md_avg = der_avg = 0
    for r in range(R):
        fa_r = fa[r] / nframes[r]                  # int64 tensor / int -> float32 tensor, as predict.py:667-668
        md_r = md[r] / nframes[r]
        fa_avg = fa_avg + fa_r
        md_avg = md_avg + md_r
        der_avg = der_avg + (fa_r + md_r)
    return {"false_alarm": fa_avg / R, "missed_detection": md_avg / R, "detection_error": der_avg / R,
            "fa_frames": fa, "md_frames": md, "nframes": nframes}
