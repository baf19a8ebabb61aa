"""Oracle: threshold + median smoothing + run-length segment extraction.

TEST INFRASTRUCTURE ONLY.  Follows
  src/utils/helper.py:66-97                      median_filter (minus the hard-coded .to("cuda"))
  src/scripts/predict.py:447-458                 per-recording slicing of the flat prediction stream
  src/scripts/predict.py:472-490                 run-length encoding, fbank / SSL time base
  src/scripts/predict_sincnet.py:348-370,492-504 run-length encoding, SincNet time base
  src/scripts/predict.py:614-647                 merge_intervals_with_buffer, split_into_windows
  src/scripts/predict.py:654-673                 get_binary_tensor / false alarm / missed detection
PINNED: every function here is checked against the output of the reference's own code on seeded inputs
(tests/golden/make_reference_golden.py -> tests/golden/reference_golden.npz, tests/test_reference_golden.py).
"""

from __future__ import annotations

import math

import numpy as np
import torch
from scipy.signal import medfilt

from .receptive_field import get_num_frames


def median_window(SPEECH_WINDOW=0.5, window=0.02) -> int:
    """helper.py:85-87: int(0.5 / window), made odd by subtracting 1."""
    k = int(SPEECH_WINDOW / window)
    if k % 2 == 0:
        k -= 1
    return k


def median_filter(x: torch.Tensor, SPEECH_WINDOW=0.5, window=0.02) -> torch.Tensor:
    """helper.py:66-97.  (B, T) float -> (B, T) int64 0/1, zero-padded scipy medfilt per row."""
    k = median_window(SPEECH_WINDOW, window)
    x = torch.where(x < 0.5, 0, 1).cpu()
    for i in range(len(x)):
        x[i] = torch.from_numpy(medfilt(x[i].numpy(), kernel_size=k))
    return x


def slice_recordings(preds_flat, durations, frame_shift=0.01, sincnet=False):
    """predict.py:447-458 / predict_sincnet.py:330-337: consecutive slices of the flat stream."""
    out, start, end = [], 0, 0
    for d in durations:
        start = end
        if sincnet:
            end = start + math.ceil(get_num_frames(16000 * d)) + 1
        else:
            end = start + math.ceil(d / frame_shift) + 1
        end = min(end, len(preds_flat))
        out.append(preds_flat[start:end])
    return out


def rle_segments(stream, frame_shift=0.01):
    """predict.py:472-490: the reference's per-frame Python loop, verbatim semantics."""
    pred_intervals = []
    start = None
    n = len(stream)
    for k in range(n):
        value = stream[k]
        if value >= 0.5:
            if start is None:
                start = k * frame_shift
        else:
            if start is not None:
                end = (k - 1) * frame_shift
                start = round(start, 2)
                end = round(end, 2)
                if end - start > 0.0:
                    pred_intervals.append((start, end))
                start = None
    if start is not None:
        end = (n - 1) * frame_shift
        start, end = round(start, 2), round(end, 2)
        if end - start > 0.0:
            pred_intervals.append((start, end))
    return pred_intervals


def sincnet_timestamp(start, end, duration):
    """predict_sincnet.py:492-504 (rounds to WHOLE seconds, as the reference does)."""
    RECEPTIVE_FIELD_1, RECEPTIVE_FIELD_2 = 991, 1261
    STEP = RECEPTIVE_FIELD_2 - RECEPTIVE_FIELD_1
    HALF_DURATION = round(0.5 * RECEPTIVE_FIELD_1)
    start_time = round((start * STEP + HALF_DURATION) / 16000)
    end_time = round((end * STEP + HALF_DURATION) / 16000)
    return max(start_time, 0), min(end_time, duration)


def rle_segments_sincnet(stream, duration):
    """predict_sincnet.py:348-370."""
    pred_intervals = []
    start = None
    n = len(stream)
    for k in range(n):
        if stream[k] >= 0.5:
            if start is None:
                start = k
        else:
            if start is not None:
                s, e = sincnet_timestamp(start, k - 1, duration)
                if e - s > 0.0:
                    pred_intervals.append((s, e))
                start = None
    if start is not None:
        s, e = sincnet_timestamp(start, n - 1, duration)
        if e - s > 0.0:
            pred_intervals.append((s, e))
    return pred_intervals


def rle_frames(stream):
    """Integer-frame view of the same RLE: (first, last_inclusive) of every run of >= 2 frames
    (a 1-frame run has end - start == 0 and is dropped by predict.py:481)."""
    a = np.asarray(stream) >= 0.5
    if a.size == 0:
        return []
    d = np.diff(np.concatenate([[0], a.astype(np.int8), [0]]))
    starts = np.nonzero(d == 1)[0]
    ends = np.nonzero(d == -1)[0] - 1
    return [(int(s), int(e)) for s, e in zip(starts, ends) if e > s]


def merge_intervals_with_buffer(intervals, total_duration, buffer):
    """predict.py:614-634."""
    if len(intervals) == 0:
        return []
    intervals = sorted(intervals, key=lambda x: x[0])
    widened = [[max(a - buffer, 0), min(b + buffer, total_duration)] for a, b in intervals]
    merged = []
    start, end = widened[0]
    for i in range(1, len(widened)):
        if widened[i][0] <= end:
            end = widened[i][1]
        else:
            merged.append([start, end])
            start, end = widened[i]
    merged.append([start, end])
    return merged


def split_into_windows(intervals, window=10):
    """predict.py:638-647."""
    out = []
    for start, end in intervals:
        while end - start > window:
            out.append([start, start + window])
            start += window
        if end - start > 0.1:
            out.append([start, end])
    return out


def get_binary_tensor(intervals, total_duration, frame_shift):
    """predict.py:654-663."""
    t = torch.zeros(math.ceil(total_duration / frame_shift))
    for start, end in intervals:
        t[int(start / frame_shift): int(end / frame_shift)] = 1
    return t


def get_false_alarm(gt, pred):
    """predict.py:666-668."""
    return torch.sum(torch.logical_and(gt == 0, pred == 1)) / len(gt)


def get_missed_detection(gt, pred):
    """predict.py:671-673."""
    return torch.sum(torch.logical_and(gt == 1, pred == 0)) / len(gt)
