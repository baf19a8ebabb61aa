"""Corpus-scale runs (BASELINE config 4): a synthetic corpus of fixed-length utterances sharded across the GPUs of
one box by utterance id, processed in batches on each rank, with ONE exchange at the end -- the NCCL gather of the
segment lists.  The audio is generated on the device, batch by batch, as a pure function of
(seed, utterance id, sample index), so every sharding of the corpus sees the same waveforms."""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .host import shard_range
from .runtime import gather_segments


def synth_corpus(first_utt: int, rows: int, num_samples: int, seed: int = 1234, device="cuda", out: Optional[torch.Tensor] = None):
    """(rows, num_samples) float32 CUDA tensor holding utterances first_utt .. first_utt + rows - 1."""
    dev = torch.device(device)
    if out is None:
        out = torch.empty((rows, num_samples), dtype=torch.float32, device=dev)
    wav = out[:rows]
    with torch.cuda.device(dev):
        for r0 in range(0, rows, 32768):
            r1 = min(rows, r0 + 32768)
            _lib.check(_lib.lib().b200vad_synth_corpus(wav[r0:r1].data_ptr(), first_utt + r0, r1 - r0, num_samples, seed,
                                                       torch.cuda.current_stream(dev).cuda_stream), "b200vad_synth_corpus")
    return wav


@torch.no_grad()
def run_corpus(packed: torch.Tensor, num_utts: int, num_samples: int, rank: int = 0, world: int = 1, batch_rows: int = 4096,
               seed: int = 1234, num_layers: int = 4, thr: float = 0.5, kernel: int = 49, gather: bool = True):
    """Process this rank's shard of the corpus; returns (segments (S, 3) int32 with GLOBAL utterance ids -- of the whole
    corpus on every rank when ``gather`` and world > 1, else of the shard --, number of frames processed locally)."""
    dev = packed.device
    lo, hi = shard_range(num_utts, rank, world)
    buf = torch.empty((min(batch_rows, max(hi - lo, 1)), num_samples), dtype=torch.float32, device=dev)
    pend, frames = [], 0
    for b0 in range(lo, hi, batch_rows):
        rows = min(batch_rows, hi - b0)
        wav = synth_corpus(b0, rows, num_samples, seed, dev, out=buf)
        # fixed-capacity segment output: nothing synchronises inside the loop (the batches queue back to back on the stream)
        prob, dec, seg, counts, seg_off = torch.ops.b200vad.vad_pipeline_padded(wav, None, packed, num_layers, thr, kernel)
        pend.append((b0, seg, seg_off))
        frames += dec.numel()
    segs = []
    totals = torch.stack([o[-1] for _, _, o in pend]).tolist() if pend else []       # ONE host read for the whole shard
    for (b0, seg, _), n in zip(pend, totals):
        sl = seg[: int(n)].clone()
        sl[:, 0] += b0
        segs.append(sl)
    local = torch.cat(segs) if segs else torch.empty((0, 3), dtype=torch.int32, device=dev)
    if gather and world > 1:
        return gather_segments(local, row_base=0), frames
    return local, frames
