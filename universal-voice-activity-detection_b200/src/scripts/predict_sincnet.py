"""Hot-path part of src/scripts/predict_sincnet.py (the PyanNet / SincNet flow): per-recording slicing of the flat prediction
stream with the SincNet frame count (:330-337), run-length extraction in frame indices (:348-370, on the GPU here), the
whole-second time base ``get_timestamp_from_sample_boundary`` (:492-504), merge / split (:507-540), scoring on a 20 ms grid
(:547-592), the truncated-cut CutSet output (:391-467) and ``get_new_cuts`` (:294-489) with the reference's signature."""

import os

import torch

import b200vad
from b200vad.host import merge_intervals_with_buffer, sincnet_timestamp, split_into_windows  # noqa: F401
from src.scripts.predict import get_segments
from src.utils.receptive_field import get_num_frames  # noqa: F401

FRAME_SHIFT = 0.02          # the grid the reference scores this flow on (predict_sincnet.py:575)


def get_timestamp_from_sample_boundary(start, end, duration):
    """predict_sincnet.py:492-504: frame indices -> WHOLE seconds (round() without digits), clamped to [0, duration]."""
    return sincnet_timestamp(start, end, duration)


def get_binary_tensor(intervals, total_duration, device="cuda"):
    """predict_sincnet.py:547-581: 0 / 1 tensor of ceil(total_duration / 0.02) frames."""
    from src.scripts.predict import get_binary_tensor as gbt
    return gbt(intervals, total_duration, FRAME_SHIFT, device)


def get_new_cuts(dataset_name, phase, tensor_file_name, recordings_path, cuts_path, predict_output_dir, output_filename=None,
                 buffer=0, split=False, alignment_path=None, verbose=True, device="cuda"):
    """predict_sincnet.py:294-489.  The prediction tensor (batch, frames, 1) is read from ``predict_output_dir /
    tensor_file_name`` as the reference does; recording i owns ``ceil(get_num_frames(16000 * duration)) + 1`` consecutive
    frames of the flat stream; runs become whole-second intervals (GPU run-length kernel); detection error on the 20 ms
    grid against the supervisions of the i-th cut (GPU bit-mask scoring).  Then, as the reference (:391-467): one new cut
    per predicted window, truncated from the recording's cut with the overlapping supervisions folded into one, written
    as a CutSet (one MonoCut JSON object per line) to ``predict_output_dir / output_filename``; ``alignment_path`` (a JSON
    of word alignments per recording) replaces the texts by the words inside each window.  The counters of the
    reference's report are returned under ``stats`` and printed with ``verbose``."""
    import json

    from b200vad import manifests

    alignment = json.load(open(alignment_path, "r")) if alignment_path else None
    preds = torch.load(os.path.join(predict_output_dir, tensor_file_name), map_location=device)
    recordings = [obj.to_dict() for obj in manifests.load_manifest_lazy(recordings_path)]
    all_cuts = list(manifests.load_manifest_lazy(cuts_path))
    assert len(all_cuts) >= len(recordings), "one cut per recording, in manifest order (predict_sincnet.py:339-340)"
    sup_dict = {}
    for cut in all_cuts:                                   # :322-324, over ALL cuts of the manifest
        for sup in cut.supervisions:
            sup_dict[sup["id"]] = [sup["start"], sup["duration"], sup.get("text"), 0]
    durations = [obj["duration"] for obj in recordings]
    gt_intervals = [[(sup.start, sup.start + sup.duration) for sup in all_cuts[i].supervisions] for i in range(len(recordings))]
    pred_intervals = get_segments(preds, durations, FRAME_SHIFT, buffer=buffer, split=split, sincnet=True)
    r = b200vad.score.detection_error(gt_intervals, pred_intervals, durations, FRAME_SHIFT, preds.device)
    stats = {"empty_cut": 0, "in_sup": 0, "exceed_sup": 0, "in_multiple_sup": 0, "sup_set": set()}
    new_cuts = []
    for i, obj in enumerate(recordings):
        new_cuts += manifests.new_cuts_from_windows(all_cuts[i], pred_intervals[i], sup_dict, stats, alignment=alignment,
                                                    recording_id=obj["id"])
    stats["total_sup"] = len(sup_dict)
    stats["unique_sup"] = len(stats["sup_set"])
    stats["not_in_sup"] = stats["total_sup"] - stats["unique_sup"]
    del stats["sup_set"]
    out = {"detection_error": r["detection_error"], "false_alarm": r["false_alarm"], "missed_detection": r["missed_detection"],
           "fa_frames": r["fa_frames"], "md_frames": r["md_frames"], "nframes": r["nframes"], "intervals": pred_intervals,
           "cuts": new_cuts, "stats": stats}
    if output_filename is not None:
        out["output_path"] = os.path.join(predict_output_dir, output_filename)
        manifests.save_manifest(new_cuts, out["output_path"])
    if verbose:
        print(f"Dataset: {dataset_name}, Buffer: {buffer}, Phase: {phase}")
        print("\n")
        print(f"Detection Error Rate: {out['detection_error']}")
        print(f"False Alarm Rate: {out['false_alarm']}")
        print(f"Missed Detection Rate: {out['missed_detection']}")
        print("----------------")
        print(f"Total Supervisions: {stats['total_sup']}")
        print(f"Supervisions in new cuts: {stats['in_sup']}")
        print(f"Unique Supervisions in new cuts: {stats['unique_sup']}")
        print(f"Supervisions not in new cuts: {stats['not_in_sup']}")
        print(f"Supervisions exceeding new cuts: {stats['exceed_sup']}")
        print(f"Supervisions in multiple new cuts: {stats['in_multiple_sup']}")
        print(f"Empty cuts: {stats['empty_cut']}")
        print("\n")
        print("\n")
    return out
