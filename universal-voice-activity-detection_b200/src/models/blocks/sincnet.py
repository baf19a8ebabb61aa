"""Drop-in for src/models/blocks/sincnet.py:33-103.  Same constructor, same parameters / buffers
(state-dict keys ``wav_norm1d.*``, ``conv1d.0.filterbank.{low_hz_,band_hz_,window_,n_}``,
``conv1d.{1,2}.*``, ``norm1d.{0,1,2}.*``); ``forward`` runs the sm_100a kernels
(``torch.ops.b200vad.sincnet``).  ``ParamSincFB`` / ``Encoder`` only hold the parameters of
asteroid-filterbanks==0.4's classes (requirements.txt:1); the filter synthesis runs on the GPU."""

import numpy as np
import torch
import torch.nn as nn

import b200vad
from .._base import PackedOwner, PackedWeights


class ParamSincFB(nn.Module):
    def __init__(self, n_filters=80, kernel_size=251, stride=1, sample_rate=16000, min_low_hz=50, min_band_hz=50):
        super().__init__()
        if (n_filters, kernel_size, min_low_hz, min_band_hz) != (80, 251, 50, 50):
            raise NotImplementedError("kernels are built for ParamSincFB(80, 251, min_low_hz=50, min_band_hz=50)")
        self.n_filters, self.kernel_size, self.stride, self.sample_rate = n_filters, kernel_size, stride, float(sample_rate)
        self.min_low_hz, self.min_band_hz = min_low_hz, min_band_hz
        self.half_kernel = kernel_size // 2
        self.cutoff = n_filters // 2
        low_hz, high_hz = 30.0, self.sample_rate / 2 - (min_low_hz + min_band_hz)
        mel = np.linspace(2595 * np.log10(1 + low_hz / 700), 2595 * np.log10(1 + high_hz / 700), self.cutoff + 1,
                          dtype="float32")
        hz = 700 * (10 ** (mel / 2595) - 1)
        self.low_hz_ = nn.Parameter(torch.from_numpy(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.from_numpy(np.diff(hz)).view(-1, 1))
        self.register_buffer("window_", torch.from_numpy(np.hamming(kernel_size)[: self.half_kernel]).float())
        self.register_buffer("n_", 2 * np.pi * (torch.arange(-self.half_kernel, 0.0).view(1, -1) / self.sample_rate))


class Encoder(nn.Module):
    def __init__(self, filterbank):
        super().__init__()
        self.filterbank = filterbank


class SincNet(PackedOwner, nn.Module):
    def __init__(self, sample_rate: int = 16000, stride: int = 1):
        super().__init__()
        if sample_rate != 16000:
            raise NotImplementedError("Only 16kHz audio supported for now.")
        self.stride = stride
        self.wav_norm1d = nn.InstanceNorm1d(1, affine=True)
        self.conv1d = nn.ModuleList()
        self.pool1d = nn.ModuleList()
        self.norm1d = nn.ModuleList()
        self.conv1d.append(Encoder(ParamSincFB(80, 251, stride=self.stride, sample_rate=sample_rate,
                                               min_low_hz=50, min_band_hz=50)))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(80, affine=True))
        self.conv1d.append(nn.Conv1d(80, 60, 5, stride=1))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(60, affine=True))
        self.conv1d.append(nn.Conv1d(60, 60, 5, stride=1))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(60, affine=True))
        self._packed = PackedWeights()

    def frames_time_major(self, waveforms: torch.Tensor) -> torch.Tensor:
        """(B, 1, N) -> (B, Ts, 60): the kernels' native layout (what PyanNet feeds its LSTM)."""
        assert waveforms.shape[1] == 1, f"Only single channel is supported. You have {waveforms.shape[1]}"
        if self.stride != 10:
            raise NotImplementedError("kernels are built for stride=10 (PyanNet.SINCNET_DEFAULTS)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("b200vad implements inference only; call .eval() / torch.no_grad()")
        blob = self._packed.get(self, waveforms.device, lambda: b200vad.pack_sincnet(self.state_dict(), waveforms.device, prefix=""),
                                always=self.repack_always)
        return torch.ops.b200vad.sincnet(waveforms[:, 0, :], blob)

    def forward(self, waveforms: torch.Tensor) -> torch.Tensor:
        """waveforms (batch, channel, sample) -> (batch, feature, frames) as sincnet.py:73-103."""
        return self.frames_time_major(waveforms).transpose(1, 2)
