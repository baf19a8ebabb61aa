"""CPU tests of the oracle itself: structural constants the reference states, cross-checks against
torchaudio / scipy, the committed golden fixtures, and property tests."""
import math
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import oracle
import util


def test_structural_constants():
    # src/datasets/custom_vad.py:43-48 (293 frames / 5 s, RF 991 / step 270); data/test_data.py:23 (500 frames / 5 s)
    assert oracle.get_num_frames(80000) == 293
    assert oracle.get_num_frames(128000) == 471
    assert oracle.receptive_field_size(1) == 991
    assert oracle.receptive_field_size(2) - oracle.receptive_field_size(1) == 270
    assert oracle.num_fbank_frames(80000) == 500
    assert oracle.num_fbank_frames(128000) == 800
    assert oracle.median_window(0.5, 0.01) == 49 and oracle.median_window(0.5, 0.02) == 25


def test_mel_banks_match_torchaudio():
    ta = pytest.importorskip("torchaudio.compliance.kaldi")
    fb = ta.get_mel_banks(80, 512, 16000.0, 20.0, -400.0, 100.0, -500.0, 1.0)[0]
    fb = torch.nn.functional.pad(fb, (0, 1)).t()
    assert torch.equal(fb, oracle.kaldi_mel_banks())


def test_fbank_shapes_and_padding():
    for n in (161, 400, 12345, 16000, 80000):
        x = 0.1 * torch.randn(2, n, generator=torch.Generator().manual_seed(n))
        f = oracle.lhotse_fbank(x)
        assert f.shape == (2, (n + 80) // 160, 80)
        assert torch.isfinite(f).all()
    x = 0.1 * torch.randn(3, 16000, generator=torch.Generator().manual_seed(1))
    lens = [16000, 8000, 15999]
    f = oracle.lhotse_fbank(x, lens=lens)
    assert torch.equal(f[1, :50], oracle.lhotse_fbank(x[1:2, :8000])[0])
    assert (f[1, 50:] == oracle.fbank.LOG_EPSILON).all()


def test_fbank_fp32_close_to_fp64():
    x = util.synth_wave(2, 32000, seed=3)
    assert util.feat_err(oracle.lhotse_fbank(x), oracle.lhotse_fbank(x, dtype=torch.float64)) < 1e-3


def test_golden_fixtures(golden_dir):
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, "fbank_small.npz"))
    wav = torch.from_numpy(g["wav"])
    assert torch.equal(wav, util.synth_wave(2, 16000, seed=5))          # the generator is deterministic
    feats = oracle.lhotse_fbank(wav)
    assert util.feat_err(feats, torch.from_numpy(g["feats"])) < 1e-5
    g2 = np.load(os.path.join(golden_dir, "pyannet2_small.npz"))
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, seed=42)
    with torch.no_grad():
        p = o(torch.from_numpy(g["feats"])).squeeze(-1)
    assert util.prob_err(p, torch.from_numpy(g2["prob"])) < 1e-5
    g3 = np.load(os.path.join(golden_dir, "pyannet_small.npz"))
    o2 = util.make_oracle("PyanNet", {}, seed=42)
    with torch.no_grad():
        s = o2.model.sincnet(wav.unsqueeze(1))
    assert util.feat_err(s, torch.from_numpy(g3["sincnet"])) < 1e-4
    g4 = np.load(os.path.join(golden_dir, "postproc.npz"))
    prob = torch.from_numpy(g4["prob"])
    assert torch.equal(oracle.median_filter(prob.clone(), window=0.01), torch.from_numpy(g4["med49"]).long())
    assert torch.equal(oracle.median_filter(prob.clone(), window=0.02), torch.from_numpy(g4["med25"]).long())


def test_sinc_filters_shape_and_symmetry():
    fb = oracle.ParamSincFB(80, 251, stride=10)
    f = fb.filters()
    assert f.shape == (80, 1, 251)
    cos, sin = f[:40, 0], f[40:, 0]
    assert torch.allclose(cos, cos.flip(1)) and torch.allclose(sin, -sin.flip(1))
    assert torch.allclose(cos[:, 125], torch.ones(40)) and (sin[:, 125] == 0).all()
    assert tuple(fb.window_.shape) == (125,) and tuple(fb.n_.shape) == (1, 125)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 1), min_size=1, max_size=300), st.sampled_from([49, 25, 5]))
def test_medfilt_is_majority_vote(bits, k):
    from scipy.signal import medfilt
    x = np.array(bits, dtype=np.int64)
    want = (np.convolve(x, np.ones(k, dtype=np.int64), "same") >= (k + 1) // 2).astype(np.int64) if len(x) >= 1 else x
    # np.convolve 'same' with len(x) < k returns length k; compute the sliding sum explicitly instead
    pad = np.concatenate([np.zeros(k // 2, np.int64), x, np.zeros(k // 2, np.int64)])
    want = np.array([pad[i:i + k].sum() >= (k + 1) // 2 for i in range(len(x))], dtype=np.int64)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = medfilt(x, kernel_size=k)
    assert np.array_equal(got, want)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 1), min_size=0, max_size=400))
def test_rle_round_trip(bits):
    """segments -> mask == mask with single-frame runs removed; seconds view == integer-frame view."""
    frames = oracle.postproc.rle_frames(bits)
    mask = np.zeros(len(bits), dtype=np.int64)
    for a, b in frames:
        mask[a:b + 1] = 1
    x = np.array(bits, dtype=np.int64)
    if len(x):
        prev = np.concatenate([[0], x[:-1]])
        nxt = np.concatenate([x[1:], [0]])
        want = x * ((prev | nxt) > 0)
    else:
        want = x
    assert np.array_equal(mask, want)
    secs = oracle.rle_segments(bits, 0.01)
    assert secs == [(round(a * 0.01, 2), round(b * 0.01, 2)) for a, b in frames]


def test_merge_and_split():
    iv = [(0.5, 1.0), (0.9, 2.0), (5.0, 17.3), (30.0, 30.05)]
    assert oracle.merge_intervals_with_buffer(iv, 40.0, 0) == [[0.5, 2.0], [5.0, 17.3], [30.0, 30.05]]
    assert oracle.merge_intervals_with_buffer([], 10.0, 1.0) == []
    assert oracle.merge_intervals_with_buffer([(1.0, 2.0), (2.5, 3.0)], 3.2, 0.3) == [[0.7, 3.2]]
    out = oracle.split_into_windows([[5.0, 17.3], [30.0, 30.05]], 10)
    assert out[0] == [5.0, 15.0] and abs(out[1][1] - 17.3) < 1e-9 and len(out) == 2


def test_vadmodel_predict_step_shapes():
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80})
    f = torch.randn(3, 120, 80)
    d = o.predict_step({"inputs": f})
    assert d.shape == (3, 120, 1) and d.dtype == torch.int64
    o2 = util.make_oracle("PyanNet", {})
    p = o2.probabilities({"inputs": 0.1 * torch.randn(2, 16000)})
    assert p.shape == (2, oracle.get_num_frames(16000), 1)


@pytest.mark.parametrize("n", [16000, 12345, 80000, 128000, 400])        # torchaudio needs n >= one frame
def test_fbank_stages_match_torchaudio_kaldi(n):
    """lhotse's Fbank differs from torchaudio.compliance.kaldi.fbank only in WHERE the DC offset and the pre-emphasis are
    applied (whole signal instead of per frame).  With those two steps applied by hand and switched
    off in torchaudio, everything else -- snip_edges=False mirror framing, povey window, 512-point power spectrum, Kaldi
    mel banks, log(max(x, eps)) -- must agree: an independent implementation of those stages."""
    ta = pytest.importorskip("torchaudio.compliance.kaldi")
    g = torch.Generator().manual_seed(n)
    w = 0.1 * torch.randn(1, n, generator=g) + 0.01
    ours = oracle.lhotse_fbank(w)[0]
    x = w - w.mean(dim=1, keepdim=True)
    x = x - 0.97 * torch.cat([x[:, :1], x[:, :-1]], dim=1)
    ref = ta.fbank(x, dither=0.0, remove_dc_offset=False, preemphasis_coefficient=0.0, snip_edges=False, window_type="povey",
                   num_mel_bins=80, low_freq=20.0, high_freq=-400.0, frame_length=25.0, frame_shift=10.0, sample_frequency=16000.0,
                   use_energy=False, round_to_power_of_two=True, use_log_fbank=True, use_power=True, subtract_mean=False)
    assert ours.shape == ref.shape == ((n + 80) // 160, 80)
    assert util.feat_err(ours, ref) <= 2e-5, util.feat_err(ours, ref)
