"""Hot-path part of src/scripts/predict.py: per-recording slicing of the flat prediction stream
(:447-458), run-length segment extraction (:472-490, on the GPU here), merge_intervals_with_buffer
(:614-634), split_into_windows (:638-647), and a manifest-free ``predict_vad``.  The reference's
``predict_vad(**config)`` reads lhotse manifests and checkpoints from hard-coded paths
(:55-409, out of scope); this one takes in-memory waveforms."""

from typing import List, Optional, Sequence

import torch

import b200vad
from b200vad.host import (merge_intervals_with_buffer, recording_offsets, segments_to_intervals,  # noqa: F401
                          split_into_windows)


def get_segments(test_preds: torch.Tensor, durations: Optional[Sequence[float]] = None, frame_shift: float = 0.02,
                 buffer: float = 0, split: bool = False, sincnet: bool = False) -> List[list]:
    """test_preds: (rows, frames, 1) or (rows, frames) decisions / probabilities on the GPU (values >= 0.5
    count as speech, predict.py:473).  With ``durations`` the flat stream is re-sliced per recording as
    predict.py:447-458 does; otherwise every row is one recording.  Returns, per recording, the
    [(start_s, end_s)] list the reference builds (after merge / optional 10 s split)."""
    p = test_preds.squeeze(-1) if test_preds.dim() == 3 else test_preds
    dec = (p >= 0.5).to(torch.uint8) if p.dtype != torch.uint8 else p
    min_run = 1 if sincnet else 2
    if durations is None:
        seg, _ = torch.ops.b200vad.segments(dec.contiguous(), None, min_run)
        R = dec.shape[0]
        durs = [dec.shape[1] * frame_shift] * R
    else:
        offs = recording_offsets(durations, dec.numel(), frame_shift, sincnet=sincnet)
        seg, _ = torch.ops.b200vad.segments(dec.reshape(-1).contiguous(),
                                            torch.tensor(offs, dtype=torch.int64, device=dec.device), min_run)
        R = len(durations)
        durs = list(durations)
    per_rec = segments_to_intervals(seg.tolist(), R, frame_shift, sincnet_durations=durs if sincnet else None)
    out = []
    for i in range(R):
        merged = merge_intervals_with_buffer(per_rec[i], durs[i], buffer)
        out.append(split_into_windows(merged, window=10) if split else merged)
    return out


@torch.no_grad()
def predict_vad(model, waveforms: torch.Tensor, frame_shift: float = 0.01, max_rows: int = 4096, **kwargs):
    """waveforms (rows, samples) CUDA float32 -> (decisions (rows, T, 1) int64, per-row intervals).
    ``model`` is a ``VadModel``; for PyanNet2 the lhotse-style fbank is computed on the fly."""
    from src.features import Fbank, FbankConfig

    preds = []
    fb = Fbank(FbankConfig(device=str(waveforms.device))) if model.model_name == "PyanNet2" else None
    for b0 in range(0, waveforms.shape[0], max_rows):
        w = waveforms[b0:b0 + max_rows]
        inputs = fb.extract_batch(w, 16000) if fb is not None else w
        preds.append(model.predict_step({"inputs": inputs}, 0))
    test_preds = torch.cat(preds)
    return test_preds, get_segments(test_preds, None, frame_shift, sincnet=model.model_name == "PyanNet")
