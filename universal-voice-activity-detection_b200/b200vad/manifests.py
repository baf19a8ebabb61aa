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
