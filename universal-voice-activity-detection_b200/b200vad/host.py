"""Host-side (pure Python) logic of the hot path: frame arithmetic, the float-seconds
epilogue of the segment extraction, interval merging, rank sharding.  No device work here.

Reference semantics followed (paths relative to the reference root):
  src/utils/receptive_field.py:28-55,165-193   conv1d_num_frames / get_num_frames
  src/utils/helper.py:85-87                     median window size
  src/scripts/predict.py:447-458                per-recording slicing of the flat stream
  src/scripts/predict.py:472-490                round(k * frame_shift, 2) epilogue, `end - start > 0`
  src/scripts/predict_sincnet.py:492-504        SincNet frame -> whole-second timestamps
  src/scripts/predict.py:614-647                merge_intervals_with_buffer / split_into_windows
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

SINC_KERNELS = (251, 3, 5, 3, 5, 3)
SINC_STRIDES = (10, 3, 1, 3, 1, 3)


def conv1d_num_frames(num_samples, kernel_size=5, stride=1, padding=0, dilation=1) -> int:
    return 1 + (num_samples + 2 * padding - dilation * (kernel_size - 1) - 1) // stride


def get_num_frames(num_samples) -> int:
    n = num_samples
    for k, s in zip(SINC_KERNELS, SINC_STRIDES):
        n = conv1d_num_frames(n, kernel_size=k, stride=s)
    return int(n)


def receptive_field_size(num_frames: int = 1) -> int:
    size = num_frames
    for k, s in reversed(list(zip(SINC_KERNELS, SINC_STRIDES))):
        size = k + (size - 1) * s
    return size


def fbank_num_frames(num_samples: int) -> int:
    return (int(num_samples) + 80) // 160


def median_window(SPEECH_WINDOW: float = 0.5, window: float = 0.02) -> int:
    k = int(SPEECH_WINDOW / window)
    if k % 2 == 0:
        k -= 1
    return k


def recording_offsets(durations: Sequence[float], total_frames: int, frame_shift: float = 0.01, sincnet: bool = False):
    """Offsets (R+1) of the per-recording slices of the flat prediction stream."""
    offs, end = [0], 0
    for d in durations:
        if sincnet:
            end = end + math.ceil(get_num_frames(16000 * d)) + 1
        else:
            end = end + math.ceil(d / frame_shift) + 1
        end = min(end, total_frames)
        offs.append(end)
    return offs


def frames_to_seconds(first: int, last: int, frame_shift: float):
    """predict.py:474-481: start = round(k*fs, 2), end = round((k_end)*fs, 2); None if end - start <= 0."""
    start = round(first * frame_shift, 2)
    end = round(last * frame_shift, 2)
    if end - start > 0.0:
        return (start, end)
    return None


def sincnet_timestamp(start: int, end: int, duration):
    RECEPTIVE_FIELD_1, RECEPTIVE_FIELD_2 = 991, 1261
    STEP = RECEPTIVE_FIELD_2 - RECEPTIVE_FIELD_1
    HALF_DURATION = round(0.5 * RECEPTIVE_FIELD_1)
    start_time = round((start * STEP + HALF_DURATION) / 16000)
    end_time = round((end * STEP + HALF_DURATION) / 16000)
    return max(start_time, 0), min(end_time, duration)


def segments_to_intervals(seg_rows, num_streams: int, frame_shift: float = 0.01, sincnet_durations=None):
    """(S,3) int triples (stream, first, last) -> list (per stream) of (start_s, end_s) tuples.

    ``seg_rows`` is any iterable of 3-int rows ordered by (stream, first) (e.g. ``seg.tolist()``).
    With ``sincnet_durations`` the SincNet time base of predict_sincnet.py:492-504 is used; since that
    maps runs of a single frame to zero-length intervals as well, pass segments extracted with
    ``min_run=1`` for that mode.
    """
    out: List[List[Tuple[float, float]]] = [[] for _ in range(num_streams)]
    for r, a, b in seg_rows:
        if sincnet_durations is not None:
            s, e = sincnet_timestamp(a, b, sincnet_durations[r])
            if e - s > 0.0:
                out[r].append((s, e))
        else:
            iv = frames_to_seconds(a, b, frame_shift)
            if iv is not None:
                out[r].append(iv)
    return out


def merge_intervals_with_buffer(intervals, total_duration, buffer):
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
    new_intervals = []
    for start, end in intervals:
        while end - start > window:
            new_intervals.append([start, start + window])
            start += window
        if end - start > 0.1:
            new_intervals.append([start, end])
    return new_intervals


def shard_range(num_units: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `num_units` owned by `rank` (ceil(U*r/W))."""
    lo = -(-num_units * rank // world_size)
    hi = -(-num_units * (rank + 1) // world_size)
    return lo, hi


def cut_into_windows(num_samples: int, window: int = 80000, min_keep: int = 48000):
    """Reference long-form semantics (src/datasets/ami/utils.py:107,163): non-overlapping 5 s
    windows, windows of <= 3 s dropped, the rest padded to 5 s.  Returns [(start, length)]."""
    out = []
    for s in range(0, num_samples, window):
        n = min(window, num_samples - s)
        if n > min_keep:
            out.append((s, n))
    return out
