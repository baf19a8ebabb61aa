"""The oracle and the host-side drop-ins against fixtures produced by the REFERENCE'S OWN CODE.

tests/golden/reference_golden.npz / .json were written by tests/golden/make_reference_golden.py, which imports the
reference's modules from /root/reference (third-party packages stubbed where they carry no arithmetic) and runs them
on seeded inputs.  These tests pin the CPU oracle -- and through it every GPU parity test -- to the reference for
PyanNet2 / VadModel (a2, a6), median_filter (a7), the RLE and its seconds arithmetic (a8, a8'), merge / split (a10),
frame arithmetic (a5), config (a12) and the DER helpers (f1).  lhotse's Fbank (a1) and asteroid's sinc filters (a3 / a4)
stay unpinned: their arithmetic is not under /root/reference.
"""
import hashlib
import json
import os
import warnings

import numpy as np
import pytest
import torch

import oracle
import util

# the reference ran on this image's CPU; another host's BLAS may differ in the last bits
ATOL = 2e-6


@pytest.fixture(scope="module")
def ref(golden_dir):
    z = np.load(os.path.join(golden_dir, "reference_golden.npz"))
    meta = json.load(open(os.path.join(golden_dir, "reference_golden.json")))
    return z, meta


def state_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def test_seeded_construction_matches_reference(ref):
    """Same modules created in the same order: under manual_seed(42) the oracle's weights ARE the reference's."""
    z, meta = ref
    if meta["torch_version"] != torch.__version__:
        pytest.skip("fixtures were generated with another torch version (RNG stream not comparable)")
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, seed=42)
    assert {k: list(v.shape) for k, v in o.state_dict().items()} == meta["pyannet2_d80_state_keys"]
    assert state_hash(o.state_dict()) == meta["pyannet2_d80_seed42_state_sha256"]
    o768 = util.make_oracle("PyanNet2", {"encoding_dim": 768}, seed=42)
    assert state_hash(o768.state_dict()) == meta["pyannet2_d768_seed42_state_sha256"]


def test_pyannet2_forward_and_predict_step(ref):
    z, meta = ref
    if meta["torch_version"] != torch.__version__:
        pytest.skip("fixtures were generated with another torch version")
    feats = torch.from_numpy(z["d80_feats"])
    batch = {"inputs": feats, "is_voice": torch.from_numpy(z["d80_labels"])}
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, seed=42)
    with torch.no_grad():
        p = o(feats)
        assert p.shape == (3, 120, 1)
        assert np.abs(p.numpy() - z["d80_prob"]).max() <= ATOL
        d = o.predict_step(batch)
        assert d.dtype == torch.int64 and np.array_equal(d.numpy(), z["d80_predict"])
        o.model.classifier.weight.copy_(torch.from_numpy(z["d80_spread_cls_w"]))
        o.model.classifier.bias.copy_(torch.from_numpy(z["d80_spread_cls_b"]))
        ps = o(feats)
        assert np.abs(ps.numpy() - z["d80_spread_prob"]).max() <= ATOL
        ds = o.predict_step(batch).numpy()
        near = np.abs(z["d80_spread_prob"][..., 0] - 0.5) <= 1e-5
        assert 0.2 < z["d80_spread_predict"].mean() < 0.8          # non-degenerate decisions
        if not near.any():
            assert np.array_equal(ds, z["d80_spread_predict"])
    o768 = util.make_oracle("PyanNet2", {"encoding_dim": 768}, seed=42)
    with torch.no_grad():
        o768.model.classifier.weight.copy_(torch.from_numpy(z["d768_cls_w"]))
        o768.model.classifier.bias.copy_(torch.from_numpy(z["d768_cls_b"]))
        x = torch.from_numpy(z["d768_x"])
        assert np.abs(o768(x).numpy() - z["d768_prob"]).max() <= ATOL
        # encoding_dim 768 -> 20 ms frames -> median window 25 (vad_engine.py:207)
        assert np.array_equal(o768.predict_step({"inputs": x}).numpy(), z["d768_predict"])


@pytest.mark.parametrize("tag,mono", [("tiny_mono", True), ("tiny_split", False)])
def test_small_pyannet2_with_stored_weights(ref, tag, mono):
    """Independent of any RNG: every weight of a small reference PyanNet2 is in the fixture."""
    z, _ = ref
    m = oracle.models.PyanNet2(lstm={"hidden_size": 8, "num_layers": 2, "monolithic": mono}, linear={"hidden_size": 6, "num_layers": 2},
                               encoding_dim=5)
    m.build()
    m.eval()
    sd = {k[len(tag) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + "_w_")}
    assert set(sd) == set(m.state_dict())
    m.load_state_dict(sd)
    with torch.no_grad():
        p = m(torch.from_numpy(z[f"{tag}_x"]))
    assert np.abs(p.numpy() - z[f"{tag}_prob"]).max() <= ATOL


def test_median_filter(ref):
    z, _ = ref
    prob = torch.from_numpy(z["mf_prob"])
    for key, kw in (("mf_49", {"window": 0.01}), ("mf_25", {"window": 0.02}), ("mf_default", {})):
        got = oracle.median_filter(prob.clone(), **kw)
        assert got.dtype == torch.int64 and np.array_equal(got.numpy(), z[key])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = oracle.median_filter(torch.from_numpy(z["mf_short_prob"]).clone(), window=0.01)
    assert np.array_equal(got.numpy(), z["mf_short_49"])


def test_frame_arithmetic(ref):
    import b200vad.host as host
    from src.utils import receptive_field as drop_rf
    z, _ = ref
    for n, f in zip(z["rf_num_samples"].tolist(), z["rf_num_frames"].tolist()):
        if f < 1:
            continue            # shorter than one receptive field: the reference's formula goes negative, nobody calls it there
        assert oracle.get_num_frames(n) == f and host.get_num_frames(n) == f and drop_rf.get_num_frames(n) == f
    for k, s in zip((1, 2, 3, 10, 293), z["rf_field_size"].tolist()):
        assert oracle.receptive_field_size(k) == s and host.receptive_field_size(k) == s and drop_rf.receptive_field_size(k) == s
    want = z["rf_conv1d"].tolist()
    got = [oracle.conv1d_num_frames(n, 5, 1) for n in (5, 6, 100)] + [oracle.conv1d_num_frames(n, 251, 10) for n in (251, 1000)]
    assert got == want
    assert [host.conv1d_num_frames(n, 5, 1) for n in (5, 6, 100)] + [host.conv1d_num_frames(n, 251, 10) for n in (251, 1000)] == want


def test_rle_and_seconds_arithmetic(ref):
    import b200vad.host as host
    z, _ = ref
    for i in range(6):
        for tag, fs in (("10ms", 0.01), ("20ms", 0.02)):
            s = z[f"rle{i}_{tag}_stream"]
            want = [tuple(r) for r in z[f"rle{i}_{tag}_intervals"].tolist()]
            assert oracle.rle_segments(s.tolist(), fs) == want
            # integer runs (what the CUDA kernel emits) + the host epilogue reproduce the reference's floats exactly
            rows = [(0, a, b) for a, b in oracle.postproc.rle_frames(s)]
            assert host.segments_to_intervals(rows, 1, fs)[0] == want
    assert len(z["rle4_10ms_intervals"]) == 2                     # runs of 2 and 3 frames kept, single frames dropped


def test_merge_split_der_sincnet_time(ref):
    import b200vad.host as host
    z, _ = ref
    iv = [tuple(r) for r in z["merge_in"].tolist()]
    for b in (0, 0.25, 1.0):
        want = z[f"merge_b{b}"].tolist()
        assert oracle.postproc.merge_intervals_with_buffer(list(iv), 60.0, b) == want
        assert host.merge_intervals_with_buffer(list(iv), 60.0, b) == want
    sp_in = [[0.0, 35.05], [40.0, 40.05], [50.0, 60.1], [70.0, 80.0]]
    assert oracle.postproc.split_into_windows([list(x) for x in sp_in], window=10) == z["split10"].tolist()
    assert host.split_into_windows([list(x) for x in sp_in], window=10) == z["split10"].tolist()
    gt = oracle.postproc.get_binary_tensor([(0.3, 1.1), (5.2, 7.9)], 10.0, 0.01)
    pr = oracle.postproc.get_binary_tensor([(0.5, 1.5), (5.0, 7.0), (9.0, 9.5)], 10.0, 0.01)
    assert np.array_equal(np.asarray(gt), z["der_gt"]) and np.array_equal(np.asarray(pr), z["der_pred"])
    fa, md = float(oracle.postproc.get_false_alarm(gt, pr)), float(oracle.postproc.get_missed_detection(gt, pr))
    assert abs(fa - z["der_fa_md"][0]) < 1e-7 and abs(md - z["der_fa_md"][1]) < 1e-7
    for (a, b, d), want in zip(z["sinc_ts_in"].tolist(), z["sinc_ts_out"].tolist()):
        assert list(oracle.postproc.sincnet_timestamp(a, b, d)) == want
        assert list(host.sincnet_timestamp(a, b, d)) == want


def test_config_keys(ref):
    from config.config import load_config
    _, meta = ref
    want = meta["config"]
    cfg = load_config(want["feature_extractor"])
    for k, v in want.items():
        got = dict(cfg[k]) if k == "model_dict" else cfg[k]
        assert got == v, (k, got, v)
    assert load_config("fbank").frame_shift == 0.01 and load_config("fbank").model_dict.encoding_dim == 80
    assert load_config("sincnet").model_name == "PyanNet" and load_config("sincnet").model_dict.encoding_dim == 60


def _report_value(report, name):
    line = next(l for l in report.splitlines() if l.startswith(name))
    return float(line.split(":")[1].strip().replace("tensor(", "").rstrip(")"))


def test_get_new_cuts_chain(ref, golden_dir):
    """The reference's whole get_new_cuts (predict.py:412-612) ran on tests/golden/manifests/*.jsonl.gz; the oracle's
    slice -> RLE -> merge -> split -> binary tensors -> FA / MD chain and the manifest reader reproduce it."""
    from b200vad import manifests
    z, meta = ref
    g = meta["get_new_cuts"]
    recs = manifests.load_manifest(os.path.join(golden_dir, "manifests", "recordings.jsonl.gz"))
    cuts = manifests.load_manifest(os.path.join(golden_dir, "manifests", "cuts.jsonl.gz"))
    durs = [r.to_dict()["duration"] for r in recs]
    assert durs == g["durations"] and [c.recording.id for c in cuts] == [r.id for r in recs]
    preds = torch.from_numpy(z["gnc_preds"].astype(np.int64))
    streams = oracle.slice_recordings(preds.reshape(-1), durs, 0.01)
    for tag in ("b0", "b025_split"):
        want = g[tag]
        fa_avg = md_avg = 0
        for i, s in enumerate(streams):
            iv = oracle.merge_intervals_with_buffer(oracle.rle_segments(s.tolist(), 0.01), durs[i], want["buffer"])
            if want["split"]:
                iv = oracle.split_into_windows(iv, window=10)
            assert [list(x) for x in iv] == want["intervals"][i], (tag, i)
            gt = oracle.get_binary_tensor([(sup.start, sup.start + sup.duration) for sup in cuts[i].supervisions], durs[i], 0.01)
            pr = oracle.get_binary_tensor(iv, durs[i], 0.01)
            fa, md = oracle.get_false_alarm(gt, pr), oracle.get_missed_detection(gt, pr)
            assert float(fa) == want["fa"][i] and float(md) == want["md"][i], (tag, i)
            fa_avg, md_avg = fa_avg + fa, md_avg + md
        assert abs(float(fa_avg / len(durs)) - _report_value(want["report"], "False Alarm Rate")) < 1e-7
        assert abs(float(md_avg / len(durs)) - _report_value(want["report"], "Missed Detection Rate")) < 1e-7


def test_manifest_round_trip(tmp_path, golden_dir):
    from b200vad import manifests
    recs = manifests.load_manifest(os.path.join(golden_dir, "manifests", "recordings.jsonl.gz"))
    cuts = manifests.load_manifest(os.path.join(golden_dir, "manifests", "cuts.jsonl.gz"))
    for name, items in (("r.jsonl.gz", recs), ("c.jsonl", cuts)):
        p = str(tmp_path / name)
        assert manifests.save_manifest(items, p) == len(items)
        assert [x.to_dict() for x in manifests.load_manifest_lazy(p)] == [x.to_dict() for x in items]
    c = cuts[1]
    assert c.type == "MonoCut" and c.supervisions[0].recording_id == c.recording.id and c.supervisions[0].end > c.supervisions[0].start
    # builders produce the same line layout the reference's manifests have
    r = manifests.recording("rec0", 79840)
    assert r.to_dict() == recs[0].to_dict()
    sups = manifests.intervals_to_supervisions(["a", "b"], [[(0.5, 1.25)], [(0.0, 2.0), (3.0, 3.5)]])
    assert [s.id for s in sups] == ["a-vad-0", "b-vad-0", "b-vad-1"] and sups[0].duration == 0.75 and sups[2].recording_id == "b"
    cut = manifests.mono_cut("rec0-0", r, sups[:1])
    assert cut.duration == r.duration and cut.supervisions[0].id == "a-vad-0"
    # PCM WAV sources
    x = (np.sin(np.arange(16000) * 0.05) * 12000).astype(np.int16)
    wav = str(tmp_path / "audio" / "rec9.wav")
    manifests.write_wav_pcm16(wav, x)
    rec = manifests.recording("rec9", len(x), source="audio/rec9.wav")
    assert np.array_equal(manifests.load_recording_pcm16(rec, root=str(tmp_path)), x)
    with pytest.raises(ValueError):
        manifests.load_recording_pcm16(dict(rec, sources=[{"type": "url", "source": "x"}]))


def test_get_new_cuts_sincnet_chain(ref, golden_dir):
    """predict_sincnet.get_new_cuts (predict_sincnet.py:294-489), run in full by the reference's own code: SincNet frame counts per
    recording, frame-index RLE, whole-second timestamps, scoring on the 20 ms grid."""
    import b200vad.host as host
    from b200vad import manifests
    z, meta = ref
    g = meta["get_new_cuts_sincnet"]
    recs = manifests.load_manifest(os.path.join(golden_dir, "manifests", "recordings_sincnet.jsonl.gz"))
    cuts = manifests.load_manifest(os.path.join(golden_dir, "manifests", "cuts_sincnet.jsonl.gz"))
    durs = [r.duration for r in recs]
    assert durs == g["durations"]
    preds = torch.from_numpy(z["gnc_sincnet_preds"].astype(np.int64))
    streams = oracle.slice_recordings(preds.reshape(-1), durs, sincnet=True)
    offs = host.recording_offsets(durs, preds.numel(), 0.02, sincnet=True)
    assert [len(s) for s in streams] == [b - a for a, b in zip(offs[:-1], offs[1:])]
    for tag in ("b0", "b1_split"):
        want = g[tag]
        for i, s in enumerate(streams):
            iv = oracle.merge_intervals_with_buffer(oracle.rle_segments_sincnet(s.tolist(), durs[i]), durs[i], want["buffer"])
            if want["split"]:
                iv = oracle.split_into_windows(iv, window=10)
            assert [list(x) for x in iv] == want["intervals"][i], (tag, i)
            # integer runs (min_run = 1) + the host epilogue give the same intervals
            rows = [(0, a, b) for a, b in zip(*_runs(s.numpy()))]
            hv = host.merge_intervals_with_buffer(host.segments_to_intervals(rows, 1, 0.02, sincnet_durations=[durs[i]])[0], durs[i], want["buffer"])
            if want["split"]:
                hv = host.split_into_windows(hv, window=10)
            assert [list(x) for x in hv] == want["intervals"][i], (tag, i)
            gt = oracle.get_binary_tensor([(sup.start, sup.start + sup.duration) for sup in cuts[i].supervisions], durs[i], 0.02)
            pr = oracle.get_binary_tensor(iv, durs[i], 0.02)
            assert float(oracle.get_false_alarm(gt, pr)) == want["fa"][i] and float(oracle.get_missed_detection(gt, pr)) == want["md"][i]


def _runs(a):
    d = np.diff(np.concatenate([[0], (np.asarray(a) >= 0.5).astype(np.int8), [0]]))
    return np.nonzero(d == 1)[0].tolist(), (np.nonzero(d == -1)[0] - 1).tolist()


def test_loss_glue(ref):
    """binary_cross_entropy (src/utils/loss.py:56-89): the value _common_step computes and predict_step discards."""
    from src.utils.loss import binary_cross_entropy
    z, _ = ref
    bp, bt, bw = (torch.from_numpy(z[k]) for k in ("bce_pred", "bce_target", "bce_weight"))
    assert abs(float(binary_cross_entropy(bp, bt)) - float(z["bce_plain"])) <= 1e-7
    assert abs(float(binary_cross_entropy(bp, bt, weight=bw)) - float(z["bce_weighted"])) <= 1e-7


def test_receptive_field_module_functions(ref):
    """Every function of src/utils/receptive_field.py (:28-219) against the reference's own values."""
    from src.utils import receptive_field as rf
    z, _ = ref
    KS, ST, PD, DL = [251, 3, 5, 3, 5, 3], [10, 3, 1, 3, 1, 3], [0] * 6, [1] * 6
    assert [rf.multi_conv_num_frames(n, kernel_size=KS, stride=ST, padding=PD, dilation=DL) for n in (991, 16000, 80000)] == z["rf_multi_frames"].tolist()
    assert [rf.multi_conv_receptive_field_size(k, kernel_size=KS, stride=ST, dilation=DL) for k in (1, 2, 471)] == z["rf_multi_size"].tolist()
    assert [rf.conv1d_receptive_field_size(k, kernel_size=5, stride=3, dilation=1) for k in (1, 2, 10)] == z["rf_conv_size"].tolist()
    assert [rf.conv1d_receptive_field_center(f, kernel_size=251, stride=10, padding=0, dilation=1) for f in (0, 1, 100)] == z["rf_conv_center"].tolist()
    assert [rf.multi_conv_receptive_field_center(f, kernel_size=KS, stride=ST, padding=PD, dilation=DL) for f in (0, 1, 292)] == z["rf_multi_center"].tolist()


def test_pyannet_structure_against_reference_classes(ref):
    """The reference's own SincNet / PyanNet / VadModel classes, run with the oracle's restatement of asteroid's filterbank as
    their `asteroid_filterbanks` (the one absent piece): the oracle's -- and the drop-in's -- seeded weights are identical and
    the oracle reproduces SincNet output, probabilities and decisions.  Pins a3 / a4 except the filter synthesis itself."""
    from src.engines import VadModel as DropIn
    z, meta = ref
    if meta["torch_version"] != torch.__version__:
        pytest.skip("fixtures were generated with another torch version")
    o = util.make_oracle("PyanNet", {}, seed=42)
    assert {k: list(v.shape) for k, v in o.state_dict().items()} == meta["pyannet_state_keys"]
    assert state_hash(o.state_dict()) == meta["pyannet_seed42_state_sha256"]
    torch.manual_seed(42)
    d = DropIn("PyanNet", {})
    assert state_hash(d.state_dict()) == meta["pyannet_seed42_state_sha256"]
    wav = torch.from_numpy(z["pyannet_wav"])
    with torch.no_grad():
        s = o.model.sincnet(wav.unsqueeze(1))
        p = o.model(wav.unsqueeze(1))
        dec = o.predict_step({"inputs": wav})
    assert np.abs(s.numpy() - z["pyannet_sincnet"]).max() <= 1e-5
    assert np.abs(p.numpy() - z["pyannet_prob"]).max() <= ATOL
    assert np.array_equal(dec.numpy(), z["pyannet_predict"])
