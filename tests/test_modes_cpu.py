"""CPU tests of the long-form / streaming / scoring host logic and their oracle definitions."""
import math
import os
import sys

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import oracle
from oracle import longform as olf


@given(st.integers(1, 12), st.integers(1, 40), st.integers(0, 400))
@settings(max_examples=60, deadline=None)
def test_stitch_center_definition(W, hop, extra):
    Tw = hop + 2 * (extra % (hop + 1))              # Tw >= hop, same parity
    prob = np.arange(W * Tw, dtype=np.float32).reshape(W, Tw)
    L = (W - 1) * hop + Tw
    out = olf.stitch_center(prob, hop, L)
    margin = (Tw - hop) // 2
    for f in range(L):
        # the chosen window contains f and no other window has its centre strictly nearer (ties go to the earlier one)
        w = int(out[f]) // Tw
        assert int(out[f]) % Tw == f - w * hop
        centre = lambda i: i * hop + (Tw - 1) / 2.0
        best = min(abs(f - centre(i)) for i in range(W) if 0 <= f - i * hop < Tw)
        assert abs(f - centre(w)) <= best + hop / 2.0 + 1
        if margin <= f < (W - 1) * hop + margin:
            assert margin <= f - w * hop < margin + hop
    if Tw == hop:
        assert np.array_equal(out, prob.reshape(-1))


@given(st.integers(1, 2_000_000))
@settings(max_examples=50, deadline=None)
def test_window_cutting_matches_oracle(n):
    import b200vad
    for hop in (80000, 40000, 16000):
        lf = b200vad.LongFormVad(None, window=80000, hop=hop)
        assert lf.windows(n) == olf.cut_windows(n, 80000, hop)
    wins = olf.cut_windows(n, 80000, 80000)
    assert all(v > 48000 for _, v in wins)                          # .filter(duration > 3)
    assert len(wins) == n // 80000 + (1 if n % 80000 > 48000 else 0)


def test_interval_frames_and_rates_follow_reference_arithmetic():
    import b200vad
    fs = 0.01
    ivs = [[(0.0, 0.29), (1.005, 2.5)], [], [(0.07, 0.07), (3.0, 9.99)]]
    t = b200vad.score.intervals_to_frames(ivs, fs)
    assert t.tolist() == [[0, int(0.0 / fs), int(0.29 / fs)], [0, int(1.005 / fs), int(2.5 / fs)], [2, int(0.07 / fs), int(0.07 / fs)],
                          [2, int(3.0 / fs), int(9.99 / fs)]]
    # the oracle's binary tensors agree with a direct rasterisation of those frame pairs
    g = oracle.get_binary_tensor(ivs[0], 3.0, fs)
    want = torch.zeros(math.ceil(3.0 / fs))
    for _, a, b in t.tolist()[:2]:
        want[a:b] = 1
    assert torch.equal(g, want)
