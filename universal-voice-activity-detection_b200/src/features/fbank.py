"""lhotse-shaped feature extractor over the fused sm_100a fbank kernel.

Stands in for ``lhotse.Fbank(FbankConfig(sampling_rate=16000, device="cuda"))`` at the reference's call
sites (src/utils/helper.py:120, src/datasets/ami/utils.py:152-163 and siblings): ``extract``,
``extract_batch``, ``frame_shift``, ``feature_dim``.  Only the configuration the reference uses is
implemented (FbankConfig defaults: 25 ms / 10 ms, povey window, 80 mel bins 20..7600 Hz, dither 0,
snip_edges False, pre-emphasis 0.97, DC removal); anything else raises."""

from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

import b200vad  # noqa: F401


@dataclass
class FbankConfig:
    sampling_rate: int = 16000
    frame_length: float = 0.025
    frame_shift: float = 0.01
    round_to_power_of_two: bool = True
    remove_dc_offset: bool = True
    preemph_coeff: float = 0.97
    window_type: str = "povey"
    dither: float = 0.0
    snip_edges: bool = False
    energy_floor: float = 1.1920928955078125e-07
    raw_energy: bool = True
    use_energy: bool = False
    use_fft_mag: bool = False
    low_freq: float = 20.0
    high_freq: float = -400.0
    num_filters: int = 80
    num_mel_bins: Optional[int] = None
    norm_filters: bool = False
    device: str = "cuda"


class Fbank:
    name = "kaldi-fbank"

    def __init__(self, config: Optional[FbankConfig] = None):
        self.config = config if config is not None else FbankConfig()
        if self.config != FbankConfig(device=self.config.device):
            raise NotImplementedError("the fused kernel implements FbankConfig defaults at 16 kHz only")
        if not str(self.config.device).startswith("cuda"):
            raise b200vad.B200VadError("Fbank runs on CUDA only (there is no CPU implementation of this path)")

    @property
    def device(self):
        return torch.device(self.config.device)

    @property
    def frame_shift(self) -> float:
        return self.config.frame_shift

    def feature_dim(self, sampling_rate: int = 16000) -> int:
        return self.config.num_filters

    def _to_dev(self, s) -> torch.Tensor:
        t = torch.from_numpy(s) if isinstance(s, np.ndarray) else s
        t = t.to(self.device, torch.float32)
        if t.dim() == 2 and t.shape[0] == 1:
            t = t[0]
        return t

    def extract(self, samples, sampling_rate: int = 16000):
        assert sampling_rate == self.config.sampling_rate
        if isinstance(samples, (list, tuple)):
            return self.extract_batch(samples, sampling_rate)
        is_numpy = isinstance(samples, np.ndarray)
        t = self._to_dev(samples)
        if t.dim() == 1:
            out = torch.ops.b200vad.fbank(t.unsqueeze(0), None)[0]
        else:
            out = torch.ops.b200vad.fbank(t, None)
        return out.cpu().numpy() if is_numpy else out

    def extract_batch(self, samples: Union[torch.Tensor, Sequence], sampling_rate: int = 16000, lengths=None):
        assert sampling_rate == self.config.sampling_rate
        if isinstance(samples, torch.Tensor) and samples.dim() == 2:
            lens = None if lengths is None else torch.as_tensor(lengths, dtype=torch.int32, device=self.device)
            return torch.ops.b200vad.fbank(samples.to(self.device, torch.float32), lens)
        seqs = [self._to_dev(s) for s in samples]
        lens = torch.tensor([s.numel() for s in seqs], dtype=torch.int32, device=self.device)
        N = int(lens.max())
        batch = torch.zeros((len(seqs), N), dtype=torch.float32, device=self.device)
        for i, s in enumerate(seqs):
            batch[i, : s.numel()] = s
        feats = torch.ops.b200vad.fbank(batch, lens)
        out: List[torch.Tensor] = [feats[i, : (int(n) + 80) // 160] for i, n in enumerate(lens.tolist())]
        if isinstance(samples[0], np.ndarray):
            return [o.cpu().numpy() for o in out]
        return out
