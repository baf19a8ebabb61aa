"""Drop-in for src/models/segmentation/PyanNet2.py:60-187: feature-input frame classifier
(4 x BiLSTM(128) -> 2 x Linear(128)+LeakyReLU -> Linear(1) -> Sigmoid).  Same constructor, same
parameters and state-dict keys (torch ``nn.LSTM`` / ``nn.Linear`` own the fp32 weights); ``forward``
runs the sm_100a kernels instead of cuDNN / cuBLAS."""

import torch

from ._head import Base, HeadMixin, merge_dict


class PyanNet2(HeadMixin, Base):
    def __init__(self, lstm: dict = None, linear: dict = None, encoding_dim: int = 768, sample_rate: int = 16000,
                 num_channels: int = 1):
        super(PyanNet2, self).__init__()
        lstm = merge_dict(self.LSTM_DEFAULTS, lstm)
        lstm["batch_first"] = True
        linear = merge_dict(self.LINEAR_DEFAULTS, linear)
        self.save_hyperparameters("lstm", "linear")
        self._make_head(lstm, linear, encoding_dim)

    def forward(self, audio_feats: torch.Tensor) -> torch.Tensor:
        """audio_feats (batch, frames, features) -> scores (batch, frames, 1)."""
        return self._head_forward(audio_feats)
