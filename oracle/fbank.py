"""Oracle: lhotse ``Fbank(FbankConfig(sampling_rate=16000))`` restated on torch CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: lhotse is an
un-pinned editable checkout in the reference (requirements.txt:13) and is absent
from /root/reference; the arithmetic below restates upstream
``lhotse/features/kaldi/layers.py::Wav2Win / Wav2LogFilterBank`` and
``extractors.py::Fbank`` with ``FbankConfig`` defaults.  Every stage after the whole-signal DC removal / pre-emphasis
is cross-checked against torchaudio.compliance.kaldi.fbank (tests/test_oracle.py), the mel banks bit for bit.  Reference call sites:
src/utils/helper.py:120-130, src/datasets/ami/utils.py:152-163 (and the 17
sibling ``src/datasets/*/utils.py``).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SAMPLE_RATE = 16000
FRAME_LEN = 400        # 25 ms
FRAME_SHIFT = 160      # 10 ms
FFT_LEN = 512          # round_to_power_of_two
NUM_MEL = 80
LOW_FREQ = 20.0
HIGH_FREQ = -400.0     # relative to Nyquist -> 7600 Hz
PREEMPH = 0.97
EPSILON = 1.1920928955078125e-07          # torch.finfo(torch.float).eps
LOG_EPSILON = -23.025850929940457         # lhotse feature pad value, ln(1e-10)


def num_fbank_frames(num_samples: int) -> int:
    """snip_edges=False frame count: (N + shift//2) // shift."""
    return (int(num_samples) + FRAME_SHIFT // 2) // FRAME_SHIFT


def povey_window(dtype=torch.float32) -> torch.Tensor:
    """hann(400, periodic=False) ** 0.85 (lhotse ``create_frame_window('povey')``)."""
    return torch.hann_window(FRAME_LEN, periodic=False, dtype=dtype).pow(0.85)


def _mel(f):
    return 1127.0 * math.log(1.0 + f / 700.0)


def kaldi_mel_banks(dtype=torch.float32) -> torch.Tensor:
    """(257, 80) triangular mel weights, Kaldi / torchaudio-compatible.

    Same arithmetic as ``torchaudio.compliance.kaldi.get_mel_banks(80, 512, 16000.,
    20., -400., 100., -500., 1.0)`` (vtln off), transposed and padded with a zero
    Nyquist row as lhotse does.  A unit test cross-checks it against torchaudio.
    """
    nyquist = 0.5 * SAMPLE_RATE
    high = HIGH_FREQ + nyquist if HIGH_FREQ <= 0 else HIGH_FREQ
    num_fft_bins = FFT_LEN // 2
    fft_bin_width = SAMPLE_RATE / FFT_LEN
    mel_low, mel_high = _mel(LOW_FREQ), _mel(high)
    delta = (mel_high - mel_low) / (NUM_MEL + 1)
    b = torch.arange(NUM_MEL, dtype=dtype).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    freqs = fft_bin_width * torch.arange(num_fft_bins, dtype=dtype)
    mel = (1127.0 * (1.0 + freqs / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    fb = torch.max(torch.zeros(1, dtype=dtype), torch.min(up, down))  # (80, 256)
    fb = F.pad(fb, (0, 1), mode="constant", value=0.0)                # (80, 257)
    return fb.t().contiguous()                                         # (257, 80)


def lhotse_fbank(wav: torch.Tensor, lens=None, dtype=torch.float32) -> torch.Tensor:
    """waveform (B, N) in [-1, 1] -> log-mel (B, T, 80), T = (N + 80) // 160.

    ``lens`` (optional, per-row valid sample counts) reproduces lhotse's
    per-recording extraction: each row is processed on its own prefix (own mean,
    own mirrored edges); rows are right-padded with LOG_EPSILON to the common T.
    """
    if wav.dim() == 1:
        wav = wav.unsqueeze(0)
    if lens is not None:
        T = num_fbank_frames(wav.shape[1])
        out = torch.full((wav.shape[0], T, NUM_MEL), LOG_EPSILON, dtype=dtype)
        for i, n in enumerate([int(v) for v in lens]):
            if n <= 0:
                continue
            f = lhotse_fbank(wav[i:i + 1, :n], None, dtype)[0]
            out[i, : f.shape[0]] = f
        return out

    x = wav.to(dtype)
    B, N = x.shape
    # Wav2Win: remove_dc_offset (whole signal), then pre-emphasis (whole signal)
    x = x - x.mean(dim=1, keepdim=True)
    x_prev = F.pad(x.unsqueeze(1), (1, 0), mode="replicate").squeeze(1)[:, :-1]
    x = x - PREEMPH * x_prev
    # _get_strided_batch, snip_edges=False
    T = num_fbank_frames(N)
    if T == 0:
        return torch.zeros(B, 0, NUM_MEL, dtype=dtype)
    new_n = (T - 1) * FRAME_SHIFT + FRAME_LEN
    npad = new_n - N
    npad_left = (FRAME_LEN - FRAME_SHIFT) // 2
    npad_right = npad - npad_left
    pad_left = torch.flip(x[:, :npad_left], (1,))
    if npad_right >= 0:
        pad_right = torch.flip(x[:, -npad_right:], (1,)) if npad_right > 0 else x[:, :0]
    else:
        pad_right = x[:, :0]
    x = torch.cat([pad_left, x, pad_right], dim=1)
    if npad_right < 0:
        x = x[:, :npad_right]
    x = x.contiguous()
    frames = x.as_strided((B, T, FRAME_LEN), (x.stride(0), FRAME_SHIFT, 1))
    frames = frames * povey_window(dtype)
    frames = F.pad(frames, (0, FFT_LEN - FRAME_LEN))
    spec = torch.fft.rfft(frames, dim=-1)
    power = spec.abs() ** 2
    mel = torch.matmul(power, kaldi_mel_banks(dtype))
    return torch.max(mel, torch.tensor(EPSILON, dtype=dtype)).log()
