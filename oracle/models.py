"""Oracle: the reference's frame classifiers rebuilt from the same torch modules.

TEST INFRASTRUCTURE ONLY.  Module structure and state-dict keys follow
  src/models/blocks/sincnet.py:34-103            (SincNet)
  src/models/segmentation/PyanNet.py:66-197      (PyanNet)
  src/models/segmentation/PyanNet2.py:60-187     (PyanNet2)
  src/engines/vad_engine.py:30-42,204-211,247-278 (VadModel forward / predict_step)
PyanNet2 and VadModel are PINNED to the reference's own code (weights under a seed, forward, predict_step:
tests/golden/make_reference_golden.py, tests/test_reference_golden.py).
SincNet and PyanNet are PINNED in structure: the reference's own classes, run with this module's ParamSincFB / Encoder as their
asteroid_filterbanks, give the same seeded weights and outputs (same generator / tests).
``ParamSincFB`` / ``Encoder`` restate asteroid-filterbanks==0.4 (requirements.txt:1),
which is not under /root/reference and not installed here: PARITY UNPINNED for that
filter synthesis.
The base class is nn.Module (pytorch_lightning is absent); ``hparams`` is a plain
attribute dict, enough for the reference's ``self.hparams.lstm[...]`` reads.
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .postproc import median_filter


class _HParams(dict):
    __getattr__ = dict.__getitem__


def merge_dict(defaults: dict, custom: dict = None):
    params = dict(defaults)
    if custom is not None:
        params.update(custom)
    return params


class ParamSincFB(nn.Module):
    """asteroid-filterbanks 0.4 ``ParamSincFB(80, 251, stride, sample_rate=16000,
    min_low_hz=50, min_band_hz=50)``: 40 cos + 40 sin band-pass filters."""

    def __init__(self, n_filters=80, kernel_size=251, stride=10, sample_rate=16000,
                 min_low_hz=50, min_band_hz=50):
        super().__init__()
        assert kernel_size % 2 == 1 and n_filters % 2 == 0
        self.n_filters, self.kernel_size, self.stride = n_filters, kernel_size, stride
        self.sample_rate = float(sample_rate)
        self.min_low_hz, self.min_band_hz = min_low_hz, min_band_hz
        self.half_kernel = kernel_size // 2
        self.cutoff = n_filters // 2
        low_hz = 30.0
        high_hz = self.sample_rate / 2 - (min_low_hz + min_band_hz)
        mel = np.linspace(2595 * np.log10(1 + low_hz / 700), 2595 * np.log10(1 + high_hz / 700),
                          self.cutoff + 1, dtype="float32")
        hz = 700 * (10 ** (mel / 2595) - 1)
        self.low_hz_ = nn.Parameter(torch.from_numpy(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.from_numpy(np.diff(hz)).view(-1, 1))
        window_ = np.hamming(kernel_size)[: self.half_kernel]          # half Hamming window
        self.register_buffer("window_", torch.from_numpy(window_).float())
        n_ = 2 * np.pi * (torch.arange(-self.half_kernel, 0.0).view(1, -1) / self.sample_rate)
        self.register_buffer("n_", n_)

    def filters(self) -> torch.Tensor:
        low = self.min_low_hz + torch.abs(self.low_hz_)
        high = torch.clamp(low + self.min_band_hz + torch.abs(self.band_hz_),
                           self.min_low_hz, self.sample_rate / 2)
        band = (high - low)[:, 0]
        ft_low = torch.matmul(low, self.n_)
        ft_high = torch.matmul(high, self.n_)
        cos_filters = self._make(band, ft_low, ft_high, "cos")
        sin_filters = self._make(band, ft_low, ft_high, "sin")
        return torch.cat([cos_filters, sin_filters], dim=0).view(self.n_filters, 1, self.kernel_size)

    def _make(self, band, ft_low, ft_high, kind):
        if kind == "cos":
            bp_left = ((torch.sin(ft_high) - torch.sin(ft_low)) / (self.n_ / 2)) * self.window_
            bp_center = 2 * band.view(-1, 1)
            bp_right = torch.flip(bp_left, dims=[1])
        else:
            bp_left = ((torch.cos(ft_low) - torch.cos(ft_high)) / (self.n_ / 2)) * self.window_
            bp_center = torch.zeros_like(band.view(-1, 1))
            bp_right = -torch.flip(bp_left, dims=[1])
        bp = torch.cat([bp_left, bp_center, bp_right], dim=1)
        return bp / (2 * band[:, None])


class Encoder(nn.Module):
    """asteroid-filterbanks ``Encoder``: conv1d with the filterbank's filters."""

    def __init__(self, filterbank: ParamSincFB):
        super().__init__()
        self.filterbank = filterbank

    def forward(self, x):
        return F.conv1d(x, self.filterbank.filters(), stride=self.filterbank.stride, padding=0)


class SincNet(nn.Module):
    """src/models/blocks/sincnet.py:33-103."""

    def __init__(self, sample_rate: int = 16000, stride: int = 1):
        super().__init__()
        if sample_rate != 16000:
            raise NotImplementedError("Only 16kHz audio supported for now.")
        self.stride = stride
        self.wav_norm1d = nn.InstanceNorm1d(1, affine=True)
        self.conv1d = nn.ModuleList()
        self.pool1d = nn.ModuleList()
        self.norm1d = nn.ModuleList()
        self.conv1d.append(Encoder(ParamSincFB(80, 251, stride=stride, sample_rate=sample_rate,
                                               min_low_hz=50, min_band_hz=50)))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(80, affine=True))
        self.conv1d.append(nn.Conv1d(80, 60, 5, stride=1))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(60, affine=True))
        self.conv1d.append(nn.Conv1d(60, 60, 5, stride=1))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(60, affine=True))

    def forward(self, waveforms):
        assert waveforms.shape[1] == 1, f"Only single channel is supported. You have {waveforms.shape[1]}"
        outputs = self.wav_norm1d(waveforms)
        for c, (conv1d, pool1d, norm1d) in enumerate(zip(self.conv1d, self.pool1d, self.norm1d)):
            outputs = conv1d(outputs)
            if c == 0:
                outputs = torch.abs(outputs)
            outputs = F.leaky_relu(norm1d(pool1d(outputs)))
        return outputs


class _Head(nn.Module):
    LSTM_DEFAULTS = {"hidden_size": 128, "num_layers": 4, "bidirectional": True,
                     "monolithic": True, "dropout": 0.5}
    LINEAR_DEFAULTS = {"hidden_size": 128, "num_layers": 2}

    def _make_head(self, lstm, linear, encoding_dim):
        lstm = merge_dict(self.LSTM_DEFAULTS, lstm)
        lstm["batch_first"] = True
        linear = merge_dict(self.LINEAR_DEFAULTS, linear)
        self.hparams.update(lstm=lstm, linear=linear)
        self.encoding_dim = encoding_dim
        if lstm["monolithic"]:
            multi = dict(lstm)
            del multi["monolithic"]
            self.lstm = nn.LSTM(encoding_dim, **multi)
        else:
            nl = lstm["num_layers"]
            if nl > 1:
                self.dropout = nn.Dropout(p=lstm["dropout"])
            one = dict(lstm)
            one["num_layers"] = 1
            one["dropout"] = 0.0
            del one["monolithic"]
            width = lstm["hidden_size"] * (2 if lstm["bidirectional"] else 1)
            self.lstm = nn.ModuleList([nn.LSTM(encoding_dim if i == 0 else width, **one) for i in range(nl)])
        if linear["num_layers"] < 1:
            return
        width = lstm["hidden_size"] * (2 if lstm["bidirectional"] else 1)
        dims = [width] + [linear["hidden_size"]] * linear["num_layers"]
        self.linear = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def build(self):
        if self.hparams.linear["num_layers"] > 0:
            in_features = self.hparams.linear["hidden_size"]
        else:
            in_features = self.hparams.lstm["hidden_size"] * (2 if self.hparams.lstm["bidirectional"] else 1)
        self.classifier = nn.Linear(in_features, 1)
        self.activation = nn.Sigmoid()

    def _head_forward(self, outputs):
        if self.hparams.lstm["monolithic"]:
            outputs, _ = self.lstm(outputs)
        else:
            for i, lstm in enumerate(self.lstm):
                outputs, _ = lstm(outputs)
                if i + 1 < self.hparams.lstm["num_layers"]:
                    outputs = self.dropout(outputs)
        if self.hparams.linear["num_layers"] > 0:
            for linear in self.linear:
                outputs = F.leaky_relu(linear(outputs))
        return self.activation(self.classifier(outputs))


class PyanNet2(_Head):
    """src/models/segmentation/PyanNet2.py:69-187 -- features (B,T,D) -> (B,T,1)."""

    def __init__(self, lstm: dict = None, linear: dict = None, encoding_dim: int = 768,
                 sample_rate: int = 16000, num_channels: int = 1):
        super().__init__()
        self.hparams = _HParams()
        self._make_head(lstm, linear, encoding_dim)

    def forward(self, audio_feats):
        return self._head_forward(audio_feats)


class PyanNet(_Head):
    """src/models/segmentation/PyanNet.py:76-197 -- waveform (B,1,N) -> (B,Ts,1)."""

    SINCNET_DEFAULTS = {"stride": 10}

    def __init__(self, sincnet: dict = None, lstm: dict = None, linear: dict = None,
                 encoding_dim: int = 60, sample_rate: int = 16000, num_channels: int = 1):
        super().__init__()
        self.hparams = _HParams()
        sincnet = merge_dict(self.SINCNET_DEFAULTS, sincnet)
        sincnet["sample_rate"] = sample_rate
        self.hparams.update(sincnet=sincnet)
        self.sincnet = SincNet(**sincnet)
        self._make_head(lstm, linear, encoding_dim)

    def forward(self, waveforms):
        outputs = self.sincnet(waveforms)            # (B, feature, frames)
        outputs = outputs.transpose(1, 2)            # rearrange "b f t -> b t f" (PyanNet.py:178)
        return self._head_forward(outputs)


class VadModel(nn.Module):
    """src/engines/vad_engine.py:20-42 (ctor), :69-80 (forward), :204-211 (predict_step),
    :247-278 (_common_step; the BCE loss value is discarded by predict_step and is not
    restated here)."""

    def __init__(self, model_name: str = "PyanNet2", model_dict: dict = {}, learning_rate: float = 1e-3):
        super().__init__()
        self.model_name = model_name
        self.model = PyanNet(**model_dict) if model_name == "PyanNet" else PyanNet2(**model_dict)
        self.model.build()
        self.learning_rate = learning_rate

    def forward(self, x):
        return self.model(x)

    def probabilities(self, batch):
        if self.model_name == "PyanNet":
            return self.model(batch["inputs"].unsqueeze(1))
        return self.model(batch["inputs"])

    def predict_step(self, batch, batch_idx=0):
        y_pred = self.probabilities(batch)
        window = 0.02 if self.model.encoding_dim == 768 else 0.01
        y_pred = median_filter(y_pred.squeeze(-1), window=window)
        return y_pred.unsqueeze(-1)
