"""Drop-in for src/engines/vad_engine.py: ``VadModel`` keeps its constructor, ``forward``,
``_common_step`` and ``predict_step`` (:30-42, :69-80, :204-211, :247-278).  ``predict_step`` returns
the reference's (B, T, 1) int64 0/1 tensor on the input device; forward, threshold and median filter
all stay on the GPU (no .cpu() / scipy round trip).  Training / torchmetrics logging are out of scope."""

import torch

from src.models import PyanNet, PyanNet2
from src.models._base import Base
from src.utils.helper import median_filter
from src.utils.loss import binary_cross_entropy


class VadModel(Base):
    def __init__(self, model_name: str = "PyanNet2", model_dict: dict = {}, learning_rate: float = 1e-3):
        super(VadModel, self).__init__()
        self.model_name = model_name
        self.model = PyanNet(**model_dict) if model_name == "PyanNet" else PyanNet2(**model_dict)
        self.model.build()
        self.learning_rate = learning_rate

    def forward(self, audio_feats: torch.Tensor) -> torch.Tensor:
        return self.model(audio_feats)

    def _common_step(self, batch, batch_idx):
        if self.model_name == "PyanNet":
            y_pred = self.model(batch["inputs"].unsqueeze(1))
        elif self.model_name == "PyanNet2":
            y_pred = self.model(batch["inputs"])
        y = batch.get("is_voice") if hasattr(batch, "get") else batch["is_voice"]
        if y is None:
            return {"loss": None}, y_pred, None
        loss = binary_cross_entropy(y_pred, y, weight=None)
        if torch.isnan(loss):
            return None
        return {"loss": loss}, y_pred, y

    def predict_step(self, batch, batch_idx=0):
        loss_dict, y_pred, y = self._common_step(batch, batch_idx)
        window = 0.02 if self.model.encoding_dim == 768 else 0.01
        y_pred = median_filter(y_pred.squeeze(-1), window=window)   # (batch, frames) int64
        return y_pred.unsqueeze(-1)

    def test_step(self, batch, batch_idx=0):
        """vad_engine.py:167-202 without the torchmetrics / Lightning logging: returns the stat
        scores (tp, fp, tn, fn) of the median-filtered decisions and the derived rates."""
        loss_dict, y_pred, y = self._common_step(batch, batch_idx)
        window = 0.02 if self.model.encoding_dim == 768 else 0.01
        d = median_filter(y_pred.squeeze(-1), window=window)
        # BinaryStatScores of the reference's torchmetrics objects (vad_engine.py:46-64, 186-195) as one popcount kernel
        tp, fp, tn, fn = torch.ops.b200vad.stat_scores(d.to(torch.uint8), (y.to(d.device) != 0).to(torch.uint8)).tolist()
        denom = d.shape[0] * d.shape[1]
        return {"test_detection_error_rate": (fp + fn) / denom, "test_false_alarm": fp / denom,
                "test_missed_detection": fn / denom, "stat_scores": (tp, fp, tn, fn), "test_loss": loss_dict["loss"]}

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.learning_rate)
