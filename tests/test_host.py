"""CPU tests of the host side: the C ABI library loads and exports every symbol include/b200vad.h
declares, the drop-in Python surface mirrors the reference, and the pure-host logic matches the oracle.
No compute call is made here (no GPU in this container)."""
import math
import os
import re

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import b200vad
    from b200vad import _lib
    hdr = open(os.path.join(ROOT, "include", "b200vad.h")).read()
    declared = set(re.findall(r"B200VAD_API[^;]*?\b(b200vad_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = b200vad.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b200vad.h but not exported by libb200vad.so"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.b200vad_abi_version() == 1


def test_abi_host_queries_match_reference_arithmetic():
    import b200vad
    lib = b200vad.lib()
    for n in (0, 1, 79, 80, 161, 80000, 128000, 960000):
        assert lib.b200vad_fbank_num_frames(n) == oracle.num_fbank_frames(n)
    for n in (991, 16000, 80000, 128000, 960000):
        assert lib.b200vad_sincnet_num_frames(n) == oracle.get_num_frames(n) == b200vad.host.get_num_frames(n)
    assert lib.b200vad_median_window(0.5, 0.01) == 49 and lib.b200vad_median_window(0.5, 0.02) == 25
    assert lib.b200vad_model_packed_bytes(80, 4) > 4 * 2 * (512 * 128 * 2)
    assert lib.b200vad_model_workspace_bytes(4096, 800) > 4096 * 800 * 4096


def test_no_cpu_fallback():
    import b200vad
    with pytest.raises(Exception):
        torch.ops.b200vad.fbank(torch.zeros(2, 16000), None)
    with pytest.raises(Exception):
        torch.ops.b200vad.threshold_median(torch.zeros(2, 100), 0.5, 49, True)
    from src.engines import VadModel
    m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    with pytest.raises(Exception):
        with torch.no_grad():
            m(torch.zeros(1, 10, 80))
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "universal-voice-activity-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), os.path.join(dirpath, f)


def test_dropin_surface_and_state_dict_keys():
    from config.config import load_config
    from src.engines import VadModel
    from src.models import PyanNet, PyanNet2
    cfg = load_config()
    assert cfg.frame_shift == 0.01 and cfg.model_name == "PyanNet2" and cfg.model_dict.encoding_dim == 80
    assert load_config("sincnet").model_name == "PyanNet" and load_config("wav2vec2").frame_shift == 0.02
    m = VadModel(cfg.model_name, dict(cfg.model_dict))
    o = oracle.VadModel("PyanNet2", {"encoding_dim": 80})
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert all(m.state_dict()[k].shape == v.shape for k, v in o.state_dict().items())
    m.load_state_dict(o.state_dict())
    m2, o2 = VadModel("PyanNet", {}), oracle.VadModel("PyanNet", {})
    assert list(m2.state_dict().keys()) == list(o2.state_dict().keys())
    m2.load_state_dict(o2.state_dict())
    assert "model.lstm.weight_hh_l3_reverse" in m.state_dict() and "model.sincnet.conv1d.0.filterbank.low_hz_" in m2.state_dict()
    p = PyanNet2(encoding_dim=80)
    assert p.hparams.lstm["hidden_size"] == 128 and p.hparams.linear["num_layers"] == 2 and p.hparams.lstm["batch_first"]
    q = PyanNet2(lstm={"monolithic": False}, encoding_dim=80)
    assert "lstm.3.weight_ih_l0_reverse" in q.state_dict()
    assert PyanNet().hparams.sincnet["stride"] == 10
    with pytest.raises(NotImplementedError):
        from src.models.blocks.sincnet import SincNet
        SincNet(sample_rate=8000)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 1), min_size=0, max_size=300), st.sampled_from([0.01, 0.02]))
def test_host_epilogue_matches_reference_loop(bits, fs):
    from b200vad import host
    frames = [(0, a, b) for a, b in oracle.postproc.rle_frames(bits)]
    got = host.segments_to_intervals(frames, 1, fs)[0]
    assert got == oracle.rle_segments(bits, fs)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.integers(0, 1), min_size=1, max_size=300))
def test_host_sincnet_epilogue(bits):
    from b200vad import host
    dur = len(bits) * 270 / 16000
    runs = []
    x = np.array(bits)
    d = np.diff(np.concatenate([[0], x, [0]]))
    for a, b in zip(np.nonzero(d == 1)[0], np.nonzero(d == -1)[0] - 1):
        runs.append((0, int(a), int(b)))
    assert host.segments_to_intervals(runs, 1, 0.02, sincnet_durations=[dur])[0] == oracle.rle_segments_sincnet(bits, dur)


def test_recording_offsets_and_windows():
    from b200vad import host
    durations = [4.99, 13.0, 0.5, 20.2, 7.77]
    flat = torch.arange(5000)
    offs = host.recording_offsets(durations, len(flat), 0.01)
    sl = oracle.slice_recordings(flat, durations, 0.01)
    assert [len(s) for s in sl] == [offs[i + 1] - offs[i] for i in range(5)]
    offs = host.recording_offsets([5.0, 8.0], 10_000, 0.02, sincnet=True)
    assert offs[1] == oracle.get_num_frames(80000) + 1
    assert host.cut_into_windows(960000) == [(i * 80000, 80000) for i in range(12)]
    assert host.cut_into_windows(80000 + 48000) == [(0, 80000)]          # <= 3 s remainder dropped
    assert host.cut_into_windows(80000 + 48001)[-1] == (80000, 48001)
    assert host.merge_intervals_with_buffer([(1.0, 2.0), (2.5, 3.0)], 3.2, 0.3) == oracle.merge_intervals_with_buffer([(1.0, 2.0), (2.5, 3.0)], 3.2, 0.3)
    assert host.split_into_windows([[5.0, 17.3]], 10) == oracle.split_into_windows([[5.0, 17.3]], 10)


def test_shard_range_partitions():
    from b200vad import host
    for U in (0, 1, 7, 4096, 450000):
        for W in (1, 2, 4, 8):
            parts = [host.shard_range(U, r, W) for r in range(W)]
            assert parts[0][0] == 0 and parts[-1][1] == U
            assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_receptive_field_module():
    from src.utils.receptive_field import get_num_frames, receptive_field_size, conv1d_num_frames
    assert get_num_frames(80000) == 293 and receptive_field_size(1) == 991 and receptive_field_size(2) == 1261
    assert conv1d_num_frames(80000, 251, 10) == 7975


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200vad
    lo, hi = b200vad.shard_range(10, rank, world)
    # rank r owns utterances [lo, hi); its local segments use local row ids
    seg = torch.tensor([[i, 10 * (lo + i), 10 * (lo + i) + 5] for i in range(hi - lo)] * (rank + 1), dtype=torch.int32).reshape(-1, 3)
    out = b200vad.gather_segments(seg, row_base=lo)
    q.put((rank, out.tolist()))
    dist.destroy_process_group()


def _gatherer_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200vad
    res = []
    for every in (1, 0, 2):                                    # per push / one exchange at the drain / every 2 pushes
        res.append(_gatherer_run(b200vad, rank, world, every))
    assert res[0] == res[1] == res[2]
    q.put((rank, res[0]))
    dist.destroy_process_group()


def _gatherer_run(b200vad, rank, world, every):
    g = b200vad.SegmentGatherer(device="cpu", every=every)
    for step in range(3):
        lo, hi = b200vad.shard_range(10, rank, world)
        n = (hi - lo) * (rank + 1) if step != 1 else 0       # an empty batch on every rank in the middle
        seg = torch.full((40, 3), -7, dtype=torch.int32)      # padded to a fixed capacity, garbage past the valid rows
        for i in range(n):
            seg[i] = torch.tensor([i % (hi - lo), 100 * step + i, 100 * step + i + 5], dtype=torch.int32)
        seg_off = torch.tensor([0, n], dtype=torch.int64)
        g.push(seg, seg_off, row_base=lo)
    return [o.tolist() for o in g.drain()]


def test_segment_gatherer_world2_gloo():
    """The side-stream gatherer of bench.py / corpus runs over gloo: fixed-capacity inputs, deferred completion,
    identical results on both ranks, in push order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gatherer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1] and len(res[0]) == 3
    for step, out in enumerate(res[0]):
        if step == 1:
            assert out == []
            continue
        want = [[i % 5, 100 * step + i, 100 * step + i + 5] for i in range(5)] + \
               [[5 + i % 5, 100 * step + i, 100 * step + i + 5] for i in range(10)]
        assert out == want, (step, out)


def test_gather_segments_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[i, 10 * i, 10 * i + 5] for i in range(5)] + [[i, 10 * i, 10 * i + 5] for i in range(5, 10)] * 2
    assert res[0] == res[1] == want


def test_new_cuts_from_windows_match_reference_loop():
    """The CutSet output of predict_sincnet.py:391-467 (host code): given the intervals the reference's own run produced, the
    truncated cuts, their folded supervisions and the report counters equal what the reference's loop produced
    (tests/golden/make_reference_golden.py; lhotse's MonoCut.truncate itself is restated, see b200vad/manifests.py)."""
    import json
    from b200vad import manifests
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    meta = json.load(open(os.path.join(gdir, "reference_golden.json")))["get_new_cuts_sincnet"]
    cuts = manifests.load_manifest(os.path.join(gdir, "manifests", "cuts_sincnet.jsonl.gz"))
    for tag in ("b0", "b1_split"):
        want = meta[tag]
        sup_dict = {s["id"]: [s["start"], s["duration"], s.get("text"), 0] for c in cuts for s in c["supervisions"]}
        stats = {"empty_cut": 0, "in_sup": 0, "exceed_sup": 0, "in_multiple_sup": 0, "sup_set": set()}
        got = []
        for cut, windows in zip(cuts, want["intervals"]):
            got += manifests.new_cuts_from_windows(cut, [tuple(w) for w in windows], sup_dict, stats)
        assert [json.loads(json.dumps(c)) for c in got] == want["cuts"], tag
        rep = {l.split(":")[0]: l.split(":")[1].strip() for l in want["report"].splitlines() if ":" in l}
        assert stats["in_sup"] == int(rep["Supervisions in new cuts"]) and stats["empty_cut"] == int(rep["Empty cuts"])
        assert stats["exceed_sup"] == int(rep["Supervisions exceeding new cuts"])
        assert stats["in_multiple_sup"] == int(rep["Supervisions in multiple new cuts"])
        assert len(stats["sup_set"]) == int(rep["Unique Supervisions in new cuts"])


def test_truncate_cut_semantics():
    """MonoCut.truncate as lhotse documents it: overlapping supervisions kept untrimmed and shifted, touching ones dropped,
    the span clipped to the cut, boundaries on the sample grid."""
    from b200vad import manifests
    rec = manifests.recording("r", 160000)
    sups = [manifests.supervision("a", "r", 0.5, 1.0, "A"), manifests.supervision("b", "r", 2.0, 1.0, "B"),
            manifests.supervision("c", "r", 4.0, 3.0, "C")]
    cut = manifests.mono_cut("r-0", rec, sups)
    t = manifests.truncate_cut(cut, offset=1.0, duration=3.0, new_id="x")
    assert (t["id"], t["start"], t["duration"]) == ("x", 1.0, 3.0)
    assert [(s["id"], s["start"], s["duration"]) for s in t["supervisions"]] == [("a", -0.5, 1.0), ("b", 1.0, 1.0)]   # c only touches
    t = manifests.truncate_cut(cut, offset=1.0, duration=3.0, keep_excessive_supervisions=False, new_id="x")
    assert [s["id"] for s in t["supervisions"]] == ["b"]
    t = manifests.truncate_cut(cut, offset=9.0, duration=5.0, new_id="y")
    assert t["start"] == 9.0 and t["duration"] == 1.0 and t["supervisions"] == []
    assert manifests.add_durations(0.1, 0.2, sampling_rate=16000) == 0.3
