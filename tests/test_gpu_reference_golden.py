"""GPU parity against fixtures produced by the REFERENCE'S OWN CODE (tests/golden/reference_golden.npz, written by
tests/golden/make_reference_golden.py from the modules under /root/reference): the CUDA path behind the drop-in modules
is compared with what the reference's VadModel / PyanNet2 / median_filter / RLE computed on the same seeded inputs, with
no oracle in between.  Run on the B200 box: pytest -m gpu."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import b200vad  # noqa: F401
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ref(golden_dir):
    z = np.load(os.path.join(golden_dir, "reference_golden.npz"))
    meta = json.load(open(os.path.join(golden_dir, "reference_golden.json")))
    if meta["torch_version"] != torch.__version__:
        pytest.skip("fixtures were generated with another torch version (seeded weights not reproducible)")
    return z, meta


def state_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def seeded_dropin(D, meta, dev):
    """The drop-in VadModel under manual_seed(42): its weights must be the reference's, bit for bit."""
    from src.engines import VadModel
    torch.manual_seed(42)
    m = VadModel("PyanNet2", {"encoding_dim": D}).eval()
    assert state_hash(m.state_dict()) == meta[f"pyannet2_d{D}_seed42_state_sha256"]
    return m.to(dev)


def check_decisions(dec, ref_dec, ref_prob, what):
    near = np.abs(ref_prob - 0.5) <= util.NEAR_THR
    diff = dec != ref_dec
    print(f"{what}: {int(diff.sum())} decisions differ, {int(near.sum())} of {near.size} reference frames within {util.NEAR_THR} of the threshold")
    if not near.any():
        assert not diff.any()
    else:
        # a flipped frame can move the median only inside its own window
        k = 49 if what.startswith("d80") else 25
        for b, t in zip(*np.nonzero(diff)):
            assert near[b, max(0, t - k // 2): t + k // 2 + 1].any(), (b, t)


def test_vadmodel_fbank_width(dev, ref):
    z, meta = ref
    m = seeded_dropin(80, meta, dev)
    x = torch.from_numpy(z["d80_feats"]).to(dev)
    batch = {"inputs": x, "is_voice": torch.from_numpy(z["d80_labels"]).to(dev)}
    with torch.no_grad():
        p = m(x)
        assert p.shape == (3, 120, 1)
        e = util.prob_err(p.cpu(), torch.from_numpy(z["d80_prob"]))
        d = m.predict_step(batch, 0)
        assert d.dtype == torch.int64 and d.is_cuda and np.array_equal(d.cpu().numpy(), z["d80_predict"])
        m.model.classifier.weight.copy_(torch.from_numpy(z["d80_spread_cls_w"]))
        m.model.classifier.bias.copy_(torch.from_numpy(z["d80_spread_cls_b"]))
        ps = m(x)
        es = util.prob_err(ps.cpu(), torch.from_numpy(z["d80_spread_prob"]))
        ds = m.predict_step(batch, 0)
    print(f"rel err of p vs the reference: plain {e:.2e}, spread head {es:.2e}")
    assert e <= util.PROB_RTOL and es <= util.PROB_RTOL
    check_decisions(ds.cpu().numpy()[..., 0], z["d80_spread_predict"][..., 0], z["d80_spread_prob"][..., 0], "d80 spread")


def test_vadmodel_ssl_width(dev, ref):
    z, meta = ref
    m = seeded_dropin(768, meta, dev)
    x = torch.from_numpy(z["d768_x"]).to(dev)
    with torch.no_grad():
        m.model.classifier.weight.copy_(torch.from_numpy(z["d768_cls_w"]))
        m.model.classifier.bias.copy_(torch.from_numpy(z["d768_cls_b"]))
        p = m(x)
        d = m.predict_step({"inputs": x, "is_voice": torch.zeros(2, 60, device=dev)}, 0)
    e = util.prob_err(p.cpu(), torch.from_numpy(z["d768_prob"]))
    print(f"rel err of p vs the reference (768-dim input): {e:.2e}")
    assert e <= util.PROB_RTOL
    check_decisions(d.cpu().numpy()[..., 0], z["d768_predict"][..., 0], z["d768_prob"][..., 0], "d768")


def test_median_filter_matches_reference(dev, ref):
    from src.utils.helper import median_filter
    z, _ = ref
    prob = torch.from_numpy(z["mf_prob"]).to(dev)
    for key, kw in (("mf_49", {"window": 0.01}), ("mf_25", {"window": 0.02}), ("mf_default", {})):
        got = median_filter(prob.clone(), **kw)
        assert got.dtype == torch.int64 and got.is_cuda and np.array_equal(got.cpu().numpy(), z[key]), key
    got = median_filter(torch.from_numpy(z["mf_short_prob"]).to(dev), window=0.01)
    assert np.array_equal(got.cpu().numpy(), z["mf_short_49"])


def test_segments_match_reference(dev, ref):
    from src.scripts.predict import get_segments
    import b200vad.host as host
    z, _ = ref
    for tag, fs in (("10ms", 0.01), ("20ms", 0.02)):
        rows = [z[f"rle{i}_{tag}_stream"] for i in range(6) if i != 4]
        d = torch.from_numpy(np.stack(rows)).to(dev)
        got = get_segments(d, None, fs)
        for j, i in enumerate([0, 1, 2, 3, 5]):
            want = host.merge_intervals_with_buffer([tuple(r) for r in z[f"rle{i}_{tag}_intervals"].tolist()], len(rows[j]) * fs, 0)
            assert [list(x) for x in got[j]] == [list(x) for x in want], (tag, i)
    d = torch.from_numpy(z["rle4_10ms_stream"][None]).to(dev)
    got = get_segments(d, None, 0.01)
    want = host.merge_intervals_with_buffer([tuple(r) for r in z["rle4_10ms_intervals"].tolist()], 0.12, 0)
    assert [list(x) for x in got[0]] == [list(x) for x in want]


def test_get_new_cuts_matches_reference(dev, ref, golden_dir, tmp_path):
    """The drop-in get_new_cuts (manifests in, GPU RLE + bit-mask scoring, report out) against the reference's own run of
    predict.py:412-612 on the same manifests and prediction stream."""
    from src.scripts.predict import get_new_cuts
    from b200vad import manifests
    z, meta = ref
    g = meta["get_new_cuts"]
    preds = torch.from_numpy(z["gnc_preds"].astype(np.int64)).to(dev)
    rp, cp = os.path.join(golden_dir, "manifests", "recordings.jsonl.gz"), os.path.join(golden_dir, "manifests", "cuts.jsonl.gz")

    def report_value(report, name):
        line = next(l for l in report.splitlines() if l.startswith(name))
        return float(line.split(":")[1].strip())

    for tag in ("b0", "b025_split"):
        want = g[tag]
        out = get_new_cuts("synthetic", "test", preds, rp, cp, str(tmp_path), f"pred_{tag}.jsonl.gz", buffer=want["buffer"],
                           split=want["split"], frame_shift=0.01, verbose=False)
        assert [[list(x) for x in iv] for iv in out["intervals"]] == want["intervals"], tag
        fa = [float(out["fa_frames"][i] / out["nframes"][i]) for i in range(len(want["fa"]))]
        md = [float(out["md_frames"][i] / out["nframes"][i]) for i in range(len(want["md"]))]
        assert fa == want["fa"] and md == want["md"], tag
        for key, name in (("detection_error", "Detection Error Rate"), ("false_alarm", "False Alarm Rate"), ("missed_detection", "Missed Detection Rate")):
            assert float(out[key]) == report_value(want["report"], name), (tag, key)
        sups = manifests.load_manifest(out["output_path"])
        assert len(sups) == sum(len(iv) for iv in want["intervals"])
        assert all(s.recording_id.startswith("rec") and s.duration > 0 for s in sups)


def test_get_new_cuts_sincnet_matches_reference(dev, ref, golden_dir, tmp_path):
    """The drop-in predict_sincnet.get_new_cuts against the reference's own run (predict_sincnet.py:294-489)."""
    from src.scripts.predict_sincnet import get_new_cuts, get_timestamp_from_sample_boundary
    from b200vad import manifests
    z, meta = ref
    g = meta["get_new_cuts_sincnet"]
    torch.save(torch.from_numpy(z["gnc_sincnet_preds"].astype(np.int64)), str(tmp_path / "preds.pt"))
    rp = os.path.join(golden_dir, "manifests", "recordings_sincnet.jsonl.gz")
    cp = os.path.join(golden_dir, "manifests", "cuts_sincnet.jsonl.gz")

    def report_value(report, name):
        line = next(l for l in report.splitlines() if l.startswith(name))
        return float(line.split(":")[1].strip())

    for tag in ("b0", "b1_split"):
        want = g[tag]
        out = get_new_cuts("synthetic", "test", "preds.pt", rp, cp, str(tmp_path), f"pred_{tag}.jsonl.gz", buffer=want["buffer"],
                           split=want["split"], verbose=False)
        assert [[list(x) for x in iv] for iv in out["intervals"]] == want["intervals"], tag
        fa = [float(out["fa_frames"][i] / out["nframes"][i]) for i in range(len(want["fa"]))]
        md = [float(out["md_frames"][i] / out["nframes"][i]) for i in range(len(want["md"]))]
        assert fa == want["fa"] and md == want["md"], tag
        for key, name in (("detection_error", "Detection Error Rate"), ("false_alarm", "False Alarm Rate"), ("missed_detection", "Missed Detection Rate")):
            assert float(out[key]) == report_value(want["report"], name), (tag, key)
        # the CutSet output (:391-467): the reference's own loop produced want["cuts"] and the counters of its report
        assert [json.loads(json.dumps(c)) for c in out["cuts"]] == want["cuts"], tag
        written = [json.loads(json.dumps(c)) for c in manifests.load_manifest(out["output_path"])]
        assert written == want["cuts"], tag
        for key, name in (("total_sup", "Total Supervisions"), ("in_sup", "Supervisions in new cuts"),
                          ("unique_sup", "Unique Supervisions in new cuts"), ("not_in_sup", "Supervisions not in new cuts"),
                          ("exceed_sup", "Supervisions exceeding new cuts"), ("in_multiple_sup", "Supervisions in multiple new cuts"),
                          ("empty_cut", "Empty cuts")):
            assert out["stats"][key] == int(report_value(want["report"], name + ":")), (tag, key)
    for (a, b, d), w in zip(z["sinc_ts_in"].tolist(), z["sinc_ts_out"].tolist()):
        assert list(get_timestamp_from_sample_boundary(a, b, d)) == w


def test_pyannet_against_reference_classes(dev, ref):
    """The drop-in PyanNet on the GPU against the reference's own SincNet / PyanNet classes (run with the restated filterbank)."""
    from src.engines import VadModel
    z, meta = ref
    torch.manual_seed(42)
    m = VadModel("PyanNet", {}).eval()
    assert state_hash(m.state_dict()) == meta["pyannet_seed42_state_sha256"]
    m = m.to(dev)
    wav = torch.from_numpy(z["pyannet_wav"]).to(dev)
    with torch.no_grad():
        s = m.model.sincnet(wav.unsqueeze(1)).cpu()
        p = m.model(wav.unsqueeze(1)).cpu()
        d = m.predict_step({"inputs": wav, "is_voice": torch.zeros(2, p.shape[1], device=dev)}, 0).cpu()
    assert util.feat_err(s, torch.from_numpy(z["pyannet_sincnet"])) <= util.FEAT_RTOL
    assert util.prob_err(p, torch.from_numpy(z["pyannet_prob"])) <= util.PROB_RTOL
    assert np.array_equal(d.numpy(), z["pyannet_predict"])
