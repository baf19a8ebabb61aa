"""src/utils/loss.py:29-89 glue (plain torch): the reference evaluates BCE inside
``_common_step`` even for prediction and discards it; kept for API compatibility."""

import torch
import torch.nn.functional as F


def interpolate(target: torch.Tensor, weight: torch.Tensor = None):
    num_frames = target.shape[1]
    if weight is not None and weight.shape[1] != num_frames:
        weight = F.interpolate(weight.transpose(1, 2), size=num_frames, mode="linear", align_corners=False).transpose(1, 2)
    return weight


def binary_cross_entropy(prediction: torch.Tensor, target: torch.Tensor, weight: torch.Tensor = None) -> torch.Tensor:
    if len(target.shape) == 2:
        target = target.unsqueeze(dim=2)
    if weight is None:
        return F.binary_cross_entropy(prediction, target.float())
    weight = interpolate(target, weight=weight)
    return F.binary_cross_entropy(prediction, target.float(), weight=weight.expand(target.shape))
