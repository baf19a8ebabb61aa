"""Drop-in for the hot-path part of src/utils/helper.py: ``median_filter`` (:66-97) runs on the
GPU (threshold -> sliding popcount) instead of .cpu() + scipy.medfilt per row + .to("cuda");
``pairwise`` / ``merge_dict`` (:16-29) are the small helpers the model constructors use."""

import itertools
from typing import Iterable

import torch

import b200vad  # noqa: F401  (registers torch.ops.b200vad)
from b200vad.host import median_window


def pairwise(iterable: Iterable):
    a, b = itertools.tee(iterable)
    next(b, None)
    return zip(a, b)


def merge_dict(defaults: dict, custom: dict = None):
    params = dict(defaults)
    if custom is not None:
        params.update(custom)
    return params


def median_filter(x, SPEECH_WINDOW=0.5, window=0.02):
    """(B, T) float probabilities -> (B, T) int64 0/1 on ``x.device`` (which must be CUDA):
    torch.where(x < 0.5, 0, 1) followed by a zero-padded median of int(SPEECH_WINDOW / window)
    frames (made odd), exactly helper.py:85-97."""
    k = median_window(SPEECH_WINDOW, window)
    squeeze = x.dim() == 1
    if squeeze:
        x = x.unsqueeze(0)
    y = torch.ops.b200vad.threshold_median(x.float(), 0.5, k, True)
    return y[0] if squeeze else y
