"""lhotse-style manifests without lhotse: the on-disk step either side of the hot path.

The reference reads recordings and cuts with ``lhotse.load_manifest_lazy`` (src/scripts/predict.py:434-435: one JSON
object per line of a ``.jsonl.gz`` file) and uses, on the prediction path, ``recording.to_dict()["id" / "duration"]``
(:447-449) and ``cut.supervisions[i].start / .duration / .text / .id`` (:441-444, 468-470).  lhotse is not installed
here; this module reads and writes the same line format (the dict layout of lhotse's ``Recording``, ``MonoCut`` and
``SupervisionSegment`` ``to_dict()``), keeps unknown keys, and adds a PCM WAV reader for ``sources`` of type "file" so
that a recordings manifest can feed the 16-bit PCM input of the CUDA path directly.  Pure host code.
"""

from __future__ import annotations

import gzip
import json
import os
import wave
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


class Manifest(dict):
    """One manifest line: a dict with attribute access, ``to_dict()`` and typed ``supervisions`` / ``recording`` views."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        if k == "supervisions":
            return [x if isinstance(x, Manifest) else Manifest(x) for x in v]
        if k == "recording" and isinstance(v, dict) and not isinstance(v, Manifest):
            return Manifest(v)
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self) -> Dict[str, Any]:
        return json.loads(json.dumps(self))

    @property
    def end(self) -> float:
        return self["start"] + self["duration"]


def _open(path: str, mode: str):
    return gzip.open(path, mode + "t", encoding="utf-8") if str(path).endswith(".gz") else open(path, mode, encoding="utf-8")


def load_manifest_lazy(path: str) -> Iterator[Manifest]:
    """Yield the objects of a ``.jsonl`` / ``.jsonl.gz`` manifest in file order (lhotse ``load_manifest_lazy``)."""
    with _open(path, "r") as f:
        for line in f:
            line = line.strip()
            if line:
                yield Manifest(json.loads(line))


def load_manifest(path: str) -> List[Manifest]:
    return list(load_manifest_lazy(path))


def save_manifest(items: Iterable[Dict[str, Any]], path: str) -> int:
    """Write one JSON object per line (gzip when the name ends in .gz); returns the number of lines."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    n = 0
    with _open(path, "w") as f:
        for it in items:
            f.write(json.dumps(it, ensure_ascii=False) + "\n")
            n += 1
    return n


def recording(rec_id: str, num_samples: int, sampling_rate: int = 16000, source: Optional[str] = None) -> Manifest:
    """The dict lhotse's ``Recording.to_dict()`` produces for a mono file."""
    return Manifest({"id": rec_id, "sources": [{"type": "file", "channels": [0], "source": source or f"{rec_id}.wav"}],
                     "sampling_rate": sampling_rate, "num_samples": int(num_samples),
                     "duration": num_samples / sampling_rate, "channel_ids": [0]})


def supervision(sup_id: str, recording_id: str, start: float, duration: float, text: Optional[str] = None, **custom) -> Manifest:
    d = {"id": sup_id, "recording_id": recording_id, "start": start, "duration": duration, "channel": 0}
    if text is not None:
        d["text"] = text
    if custom:
        d["custom"] = custom
    return Manifest(d)


def mono_cut(cut_id: str, rec: Dict[str, Any], supervisions: Sequence[Dict[str, Any]] = (), start: float = 0.0,
             duration: Optional[float] = None) -> Manifest:
    return Manifest({"id": cut_id, "start": start, "duration": rec["duration"] if duration is None else duration, "channel": 0,
                     "supervisions": [dict(s) for s in supervisions], "recording": dict(rec), "type": "MonoCut"})


# ---------------------------------------------------------------- cut truncation (the CutSet output of predict_sincnet.py)
# lhotse is the reference's dependency for this (requirements.txt:13, an editable checkout, no version pinned) and is absent
# here, so ``MonoCut.truncate`` is restated from its published behaviour (lhotse/cut/mono.py ``MonoCut.truncate``,
# lhotse/utils.py ``add_durations`` / ``compute_num_samples`` / ``overlaps``, lhotse/supervision.py ``with_offset``):
# parity with lhotse itself is UNPINNED; the reference's own loop around it (predict_sincnet.py:391-467) is pinned by
# tests/golden/make_reference_golden.py, which runs that loop with this function standing in for lhotse's.
def _num_samples(duration: float, sampling_rate: int) -> int:
    """lhotse ``compute_num_samples``: round(duration * sr, 8 digits) to the nearest sample, halves up."""
    from decimal import ROUND_HALF_UP, Decimal
    return int(Decimal(round(duration * sampling_rate, ndigits=8)).quantize(0, rounding=ROUND_HALF_UP))


def add_durations(*durations: float, sampling_rate: int) -> float:
    """lhotse ``add_durations``: the sum taken on the sample grid (no floating-point drift in cut boundaries)."""
    return sum(_num_samples(d, sampling_rate) for d in durations) / sampling_rate


def _overlaps(a_start: float, a_end: float, b_start: float, b_end: float) -> bool:
    """lhotse ``overlaps``: open-interval overlap; touching ends do not count."""
    from math import isclose
    return a_start < b_end and b_start < a_end and not isclose(a_start, b_end) and not isclose(b_start, a_end)


def truncate_cut(cut: Dict[str, Any], offset: float = 0.0, duration: Optional[float] = None,
                 keep_excessive_supervisions: bool = True, new_id: Optional[str] = None) -> Manifest:
    """``MonoCut.truncate(offset=, duration=, keep_excessive_supervisions=)``: a cut over [offset, offset + duration) of
    ``cut`` (clipped to its end), with the supervisions that overlap the new span (``keep_excessive_supervisions=True``) or lie
    inside it (False), shifted by -offset and NOT trimmed, sorted by start.  lhotse assigns a random uuid unless the caller
    renames the cut (the reference always does, ``.with_id``): pass ``new_id``."""
    assert offset >= 0, f"offset must be non-negative (got {offset})"
    sr = int(cut.get("recording", {}).get("sampling_rate", 16000)) if isinstance(cut.get("recording"), dict) else 16000
    new_start = max(add_durations(cut["start"], offset, sampling_rate=sr), 0)
    until = offset + (duration if duration is not None else cut["duration"])
    new_duration = add_durations(until, -offset, sampling_rate=sr)
    assert new_duration > 0.0, f"truncated cut would have non-positive duration {new_duration}"
    past_end = (new_start + new_duration) - (cut["start"] + cut["duration"])
    if past_end > 0:
        new_duration -= past_end
    sups = []
    for sup in cut.get("supervisions", []):
        sh = dict(sup)
        sh["start"] = add_durations(sup["start"], -offset, sampling_rate=sr)
        s_end = sh["start"] + sh["duration"]
        if keep_excessive_supervisions:
            keep = _overlaps(0.0, new_duration, sh["start"], s_end)
        else:   # lhotse ``overspans``: the new span covers the supervision entirely
            keep = 0.0 <= sh["start"] and s_end <= new_duration
        if keep:
            sups.append(sh)
    sups.sort(key=lambda x: x["start"])
    out = Manifest({k: v for k, v in cut.items()})
    out["id"] = new_id if new_id is not None else f"{cut['id']}-truncated"
    out["start"], out["duration"], out["supervisions"] = new_start, new_duration, sups
    return out


def new_cuts_from_windows(cut: Dict[str, Any], windows: Sequence[Tuple[float, float]], sup_dict: Dict[str, list],
                          stats: Dict[str, Any], alignment: Optional[Dict[str, Any]] = None,
                          recording_id: Optional[str] = None) -> List[Manifest]:
    """predict_sincnet.py:391-462 for one recording: one new cut per predicted window (id ``<cut id>-<j>``) carrying ONE
    supervision -- the earliest overlapping one, re-based to start 0 and stretched to the cut's duration, its text the
    space-joined texts of every overlapping supervision (or, with a word-alignment file, the words that lie inside the
    window).  Windows without supervisions are counted in ``stats['empty_cut']`` and dropped.  ``sup_dict`` maps supervision
    id -> [start, duration, text, times_seen] over the whole manifest and ``stats`` accumulates the reference's counters
    (``in_sup``, ``exceed_sup``, ``in_multiple_sup``, ``empty_cut``, ``sup_set``)."""
    out = []
    for j, (abs_start, abs_end) in enumerate(windows):
        nc = truncate_cut(cut, offset=abs_start, duration=abs_end - abs_start, keep_excessive_supervisions=True,
                          new_id=f"{cut['id']}-{j}")
        sups = nc["supervisions"]
        if len(sups) == 0:
            stats["empty_cut"] += 1
            continue
        for sup in sups:
            if sup["id"] in sup_dict:
                ent = sup_dict[sup["id"]]
                stats["in_sup"] += 1
                stats["sup_set"].add(sup["id"])
                ent[3] += 1
                cut_start, sup_start = round(nc["start"], 2), round(ent[0], 2)
                cut_end, sup_end = round(nc["start"] + nc["duration"], 2), round(ent[0] + ent[1], 2)
                if cut_start > sup_start or cut_end < sup_end:
                    stats["exceed_sup"] += 1
                if ent[3] > 1:
                    stats["in_multiple_sup"] += 1
        text = "".join(s["text"] + " " for s in sups if s.get("text") is not None).strip()
        if alignment is not None:
            words = ""
            for sup in alignment[recording_id]["supervisions"]:
                if sup["start"] + sup["duration"] >= abs_start and abs_end >= sup["start"]:       # is_overlap (:543-544)
                    for al in sup["alignment"]:
                        if sup["start"] + al["start"] >= abs_start and sup["start"] + al["end"] <= abs_end:
                            words += al["word"] + " "
                elif sup["start"] > abs_end:
                    break
            text = words.strip()
        first = dict(sups[0])
        first["start"], first["duration"], first["text"] = 0, nc["duration"], text
        nc["supervisions"] = [first]
        out.append(nc)
    return out


def intervals_to_supervisions(recording_ids: Sequence[str], intervals_per_rec: Sequence[Sequence[Tuple[float, float]]],
                              tag: str = "vad") -> List[Manifest]:
    """Predicted speech intervals -> SupervisionSegment lines (what a downstream lhotse pipeline consumes)."""
    out = []
    for rid, ivs in zip(recording_ids, intervals_per_rec):
        for j, (a, b) in enumerate(ivs):
            out.append(supervision(f"{rid}-{tag}-{j}", rid, float(a), round(float(b) - float(a), 6)))
    return out


def read_wav_pcm16(path: str) -> Tuple[np.ndarray, int]:
    """16-bit PCM WAV -> (int16 samples of channel 0, sampling rate); stdlib only."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported (sample width {w.getsampwidth()})")
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        data = np.frombuffer(w.readframes(n), dtype="<i2")
    if ch > 1:
        data = data.reshape(-1, ch)[:, 0]
    return np.ascontiguousarray(data), sr


def write_wav_pcm16(path: str, samples: np.ndarray, sampling_rate: int = 16000) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sampling_rate)
        w.writeframes(np.asarray(samples, dtype="<i2").tobytes())


def load_recording_pcm16(rec: Dict[str, Any], root: Optional[str] = None) -> np.ndarray:
    """Samples of a recordings-manifest line whose first source is a PCM WAV file (16 kHz mono expected by the path)."""
    src = rec["sources"][0]
    if src.get("type") != "file":
        raise ValueError(f"recording {rec['id']}: unsupported source type {src.get('type')!r}")
    p = src["source"]
    if root is not None and not os.path.isabs(p):
        p = os.path.join(root, p)
    data, sr = read_wav_pcm16(p)
    if sr != rec.get("sampling_rate", sr):
        raise ValueError(f"recording {rec['id']}: file rate {sr} != manifest rate {rec['sampling_rate']}")
    return data
