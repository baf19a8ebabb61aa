"""Drop-in for src/models/segmentation/PyanNet.py:66-197: SincNet front-end on the raw waveform,
then the same BiLSTM / linear / sigmoid head with encoding_dim=60."""

from typing import Optional

import torch

from ..blocks.sincnet import SincNet
from ._head import Base, HeadMixin, merge_dict


class PyanNet(HeadMixin, Base):
    SINCNET_DEFAULTS = {"stride": 10}

    def __init__(self, sincnet: Optional[dict] = None, lstm: Optional[dict] = None, linear: Optional[dict] = None,
                 encoding_dim: int = 60, sample_rate: int = 16000, num_channels: int = 1):
        super(PyanNet, self).__init__()
        sincnet = merge_dict(self.SINCNET_DEFAULTS, sincnet)
        sincnet["sample_rate"] = sample_rate
        lstm = merge_dict(self.LSTM_DEFAULTS, lstm)
        lstm["batch_first"] = True
        linear = merge_dict(self.LINEAR_DEFAULTS, linear)
        self.save_hyperparameters("sincnet", "lstm", "linear")
        self.sincnet = SincNet(**self.hparams.sincnet)
        self._make_head(lstm, linear, encoding_dim)

    def forward(self, waveforms: torch.Tensor) -> torch.Tensor:
        """waveforms (batch, channel, samples) -> scores (batch, frames, 1)."""
        outputs = self.sincnet.frames_time_major(waveforms)   # already "batch frames feature" (PyanNet.py:178)
        return self._head_forward(outputs)
