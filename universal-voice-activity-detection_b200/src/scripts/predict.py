"""Hot-path part of src/scripts/predict.py: per-recording slicing of the flat prediction stream
(:447-458), run-length segment extraction (:472-490, on the GPU here), merge_intervals_with_buffer
(:614-634), split_into_windows (:638-647), ``get_new_cuts`` (:412-612: manifests in, detection-error report out) and a
manifest-free ``predict_vad``.  The reference's ``predict_vad(**config)`` reads lhotse manifests and checkpoints from
hard-coded paths (:55-409, out of scope); this one takes in-memory waveforms."""

import os
from typing import List, Optional, Sequence

import torch

import b200vad
from b200vad.host import (merge_intervals_with_buffer, recording_offsets, segments_to_intervals,  # noqa: F401
                          split_into_windows)


def get_segments(test_preds: torch.Tensor, durations: Optional[Sequence[float]] = None, frame_shift: float = 0.02,
                 buffer: float = 0, split: bool = False, sincnet: bool = False, row_duration: Optional[float] = None) -> List[list]:
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
        if row_duration is not None:
            durs = [float(row_duration)] * R
        elif sincnet:
            # SincNet frames are 270 samples apart with a 991-sample receptive field (receptive_field.py:165-219), not frame_shift
            durs = [((dec.shape[1] - 1) * 270 + 991) / 16000.0] * R
        else:
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


def get_binary_tensor(intervals, total_duration, frame_shift, device="cuda"):
    """predict.py:654-663: 0/1 float tensor of ceil(total_duration / frame_shift) frames, on the GPU."""
    import math
    n = math.ceil(total_duration / frame_shift)
    t = torch.zeros(n, device=device)
    for start, end in intervals:
        t[int(start / frame_shift): int(end / frame_shift)] = 1
    return t


def get_false_alarm(gt_tensor, pred_tensor):
    """predict.py:666-668 (frames with gt == 0 and pred == 1, over len(gt)); counts come from the stat-score kernel."""
    tp, fp, tn, fn = torch.ops.b200vad.stat_scores((pred_tensor == 1).to(torch.uint8), (gt_tensor != 0).to(torch.uint8))
    return fp.cpu() / len(gt_tensor)


def get_missed_detection(gt_tensor, pred_tensor):
    """predict.py:671-673 (frames with gt == 1 and pred == 0, over len(gt))."""
    tp, fp, tn, fn = torch.ops.b200vad.stat_scores((pred_tensor != 0).to(torch.uint8), (gt_tensor == 1).to(torch.uint8))
    return fn.cpu() / len(gt_tensor)


def score_predictions(gt_intervals, pred_intervals, durations, frame_shift=0.01, device="cuda"):
    """The accumulation of predict.py:500-509 / 590-600 for all recordings in one pass of the bit-mask kernels.
    Returns (detection_error_rate, false_alarm_rate, missed_detection_rate) averaged over recordings."""
    r = b200vad.score.detection_error(gt_intervals, pred_intervals, durations, frame_shift, device)
    return r["detection_error"], r["false_alarm"], r["missed_detection"]


def get_new_cuts(dataset_name, phase, test_preds, recordings_path, cuts_path, predict_output_dir=None, output_filename=None,
                 buffer=0, split=False, alignment_path=None, frame_shift=0.02, verbose=True):
    """predict.py:412-612 with the reference's signature: the flat prediction stream is sliced per recording of the
    recordings manifest (:447-458), turned into intervals (RLE :472-490 on the GPU, merge :492-494, optional 10 s split
    :496-498) and scored against the supervisions of the i-th cut (:468-470, 500-509) with the bit-mask kernels; the
    report of :590-610 is printed and returned.  Manifests are lhotse ``.jsonl(.gz)`` files (b200vad.manifests).  The
    reference's writer of the new cut set is commented out (:511-588); here the predicted intervals are written as
    SupervisionSegment lines to ``predict_output_dir / output_filename`` when both are given.  ``alignment_path`` (word
    alignments for the commented-out writer) is accepted and ignored."""
    from b200vad import manifests

    recordings = [obj.to_dict() for obj in manifests.load_manifest_lazy(recordings_path)]
    all_cuts = list(manifests.load_manifest_lazy(cuts_path))
    assert len(all_cuts) >= len(recordings), "one cut per recording, in manifest order (predict.py:462-463)"
    durations = [obj["duration"] for obj in recordings]
    gt_intervals = [[(sup.start, sup.start + sup.duration) for sup in all_cuts[i].supervisions] for i in range(len(recordings))]
    pred_intervals = get_segments(test_preds, durations, frame_shift, buffer=buffer, split=split)
    r = b200vad.score.detection_error(gt_intervals, pred_intervals, durations, frame_shift, test_preds.device)
    out = {"detection_error": r["detection_error"], "false_alarm": r["false_alarm"], "missed_detection": r["missed_detection"],
           "fa_frames": r["fa_frames"], "md_frames": r["md_frames"], "nframes": r["nframes"], "intervals": pred_intervals,
           "total_supervisions": sum(len(c.supervisions) for c in all_cuts)}
    if predict_output_dir is not None and output_filename is not None:
        sups = manifests.intervals_to_supervisions([obj["id"] for obj in recordings], pred_intervals)
        out["output_path"] = os.path.join(predict_output_dir, output_filename)
        manifests.save_manifest(sups, out["output_path"])
    if verbose:
        print(f"Dataset: {dataset_name}, Buffer: {buffer}, Phase: {phase}")
        print("\n")
        print(f"Detection Error Rate: {out['detection_error']}")
        print(f"False Alarm Rate: {out['false_alarm']}")
        print(f"Missed Detection Rate: {out['missed_detection']}")
        print("----------------")
        print(f"Total Supervisions: {out['total_supervisions']}")
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
    return test_preds, get_segments(test_preds, None, frame_shift, sincnet=model.model_name == "PyanNet",
                                    row_duration=waveforms.shape[1] / 16000.0)
