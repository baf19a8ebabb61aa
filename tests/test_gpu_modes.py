"""GPU parity tests of the modes around the batch path: long-form windows + stitching (BASELINE config 3),
streaming ring buffers (config 5) and detection-error scoring."""
import math

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import b200vad
    b200vad._lib.init(0)
    return torch.device("cuda:0")


def _model(dev, feats):
    import b200vad
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    return o, b200vad.pack_model(o.model.state_dict(), dev, 80, 4)


# ---------------------------------------------------------------- long-form
@pytest.mark.parametrize("seconds", [21.0, 18.9, 17.0])      # full windows only / 3.9 s tail kept / 2 s tail dropped
def test_longform_reference_semantics(dev, seconds):
    import b200vad
    import oracle
    from oracle import longform as olf
    N = int(seconds * 16000)
    wav = util.synth_wave(1, N, seed=int(seconds * 10))[0]
    o, blob = _model(dev, oracle.lhotse_fbank(wav[:80000].unsqueeze(0)))
    ref_prob, _ = olf.longform_reference(o, oracle.lhotse_fbank, wav)
    res = b200vad.LongFormVad(blob, 4)(wav.to(dev))
    assert res["windows"] == olf.cut_windows(N)
    assert util.prob_err(res["prob"].cpu(), ref_prob) <= util.PROB_RTOL
    # integer work: oracle median filter / slicing / RLE applied to OUR probabilities must match bit for bit
    dec = oracle.median_filter(res["prob"].cpu(), window=0.01)
    stream = oracle.slice_recordings(dec.reshape(-1), [N / 16000.0], 0.01)[0]
    assert torch.equal(res["stream_dec"].cpu().long(), stream)
    want = oracle.merge_intervals_with_buffer(oracle.rle_segments(stream.tolist(), 0.01), N / 16000.0, 0)
    assert [list(x) for x in res["intervals"]] == [list(x) for x in want]


@pytest.mark.parametrize("hop", [40000, 16000])
def test_longform_overlap_stitching(dev, hop):
    import b200vad
    import oracle
    from oracle import longform as olf
    N = 16000 * 23 + 1234
    wav = util.synth_wave(1, N, seed=hop)[0]
    o, blob = _model(dev, oracle.lhotse_fbank(wav[:80000].unsqueeze(0)))
    res = b200vad.LongFormVad(blob, 4, hop=hop)(wav.to(dev))
    wins = olf.cut_windows(N, 80000, hop)
    assert res["windows"] == wins
    # window probabilities against the oracle model on the same windows (last one is ragged -> LOG_EPS padded features)
    for i in (0, len(wins) // 2, len(wins) - 1):
        s, n = wins[i]
        f = oracle.lhotse_fbank(wav[s:s + n].unsqueeze(0))[0]
        f = torch.cat([f, torch.full((500 - f.shape[0], 80), -23.025850929940457)]) if f.shape[0] < 500 else f
        with torch.no_grad():
            ref = o(f.unsqueeze(0)).squeeze()
        assert util.prob_err(res["prob"][i].cpu(), ref) <= util.PROB_RTOL
    L = (N + 80) // 160
    stitched = olf.stitch_center(res["prob"].cpu().numpy(), hop // 160, L)
    got = torch.ops.b200vad.stitch_center(res["prob"], hop // 160, L).cpu().numpy()
    assert np.array_equal(got, stitched)
    dec = oracle.median_filter(torch.from_numpy(stitched).unsqueeze(0), window=0.01)[0]
    assert torch.equal(res["stream_dec"].cpu().long(), dec)
    want = oracle.merge_intervals_with_buffer(oracle.rle_segments(dec.tolist(), 0.01), N / 16000.0, 0)
    assert [list(x) for x in res["intervals"]] == [list(x) for x in want]


def test_longform_one_hour_shapes(dev):
    """BASELINE config 3 at full size: 1 h = 720 windows x 500 frames; rows are independent, so the long-form
    result must equal the same windows pushed through the batch path, and the stream has 360 000 frames."""
    import b200vad
    torch.manual_seed(0)
    N = 3600 * 16000
    g = torch.Generator(device=dev).manual_seed(1)
    wav = 0.05 * torch.randn(N, device=dev, generator=g)
    wav[: N // 2] *= torch.sin(torch.arange(N // 2, device=dev) * (2 * math.pi * 0.3 / 16000)).abs()
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80})
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    res = b200vad.LongFormVad(blob, 4)(wav)
    assert res["prob"].shape == (720, 500) and res["stream_dec"].numel() == 360000
    p, d, _, _ = torch.ops.b200vad.vad_pipeline(wav.view(720, 80000), None, blob, 4, 0.5, 49)
    assert torch.equal(p, res["prob"]) and torch.equal(d.reshape(-1), res["stream_dec"])
    assert all(0 <= a < b <= 3600.0 for a, b in res["intervals"])
    res2 = b200vad.LongFormVad(blob, 4, hop=40000)(wav)
    assert res2["prob"].shape == (1439, 500) and res2["stream_dec"].numel() == 360000
    # interior windows of the overlapped cut that coincide with a reference window give identical rows
    assert torch.equal(res2["prob"][0::2][:719], res["prob"][:719])


# ---------------------------------------------------------------- streaming
@pytest.mark.parametrize("S,window,hop,graph", [(5, 16000, 160, True), (5, 16000, 160, False), (3, 8000, 320, True)])
def test_streaming_matches_batch_on_buffered_window(dev, S, window, hop, graph):
    import b200vad
    import oracle
    total = util.synth_wave(S, window + 40 * hop, seed=window + hop)
    o, blob = _model(dev, oracle.lhotse_fbank(total[:, :window]))
    sv = b200vad.StreamingVad(blob, 4, num_streams=S, window=window, hop=hop, use_graph=graph)
    nf = hop // 160
    fed = 0
    for step in range(window // hop + 7):
        chunk = total[:, fed:fed + hop].contiguous()
        prob, dec, ms = sv.push(chunk if step % 2 else chunk.to(dev))
        fed += hop
        if step in (0, 3, window // hop - 1, window // hop + 6):
            buf = torch.zeros(S, window)
            n = min(fed, window)
            buf[:, window - n:] = total[:, fed - n:fed]
            w, p_all, d_all = sv.snapshot()
            assert torch.equal(w, buf)                                             # ring buffer == last `window` samples
            bp, bd, _, _ = torch.ops.b200vad.vad_pipeline(buf.to(dev), None, blob, 4, 0.5, 49)
            assert torch.equal(p_all, bp.cpu()) and torch.equal(d_all, bd.cpu())   # same kernels, same rows -> identical
            assert torch.equal(prob, bp[:, -nf:].cpu()) and torch.equal(dec, bd[:, -nf:].cpu())
            with torch.no_grad():
                ref = o(oracle.lhotse_fbank(buf)).squeeze(-1)
            assert util.prob_err(p_all, ref) <= util.PROB_RTOL
            assert torch.equal(oracle.median_filter(p_all, window=0.01), d_all.long())
    sv.close()


def test_streaming_config5_shape(dev):
    """BASELINE config 5: 256 streams x 5 s window x 10 ms hop; a few pushes, per-chunk device time reported."""
    import b200vad
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80})
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    sv = b200vad.StreamingVad(blob, 4, num_streams=256, window=80000, hop=160)
    g = torch.Generator().manual_seed(0)
    times = []
    for _ in range(6):
        prob, dec, ms = sv.push(0.1 * torch.randn(256, 160, generator=g))
        times.append(ms)
    assert prob.shape == (256, 1) and dec.shape == (256, 1) and torch.isfinite(prob).all()
    w, p_all, d_all = sv.snapshot()
    assert (w[:, : 80000 - 6 * 160] == 0).all() and (w[:, 80000 - 6 * 160:] != 0).any()
    print("streaming 256 x 5 s, per-push device ms:", [round(t, 2) for t in times])
    sv.close()


# ---------------------------------------------------------------- scoring
def test_stat_scores_and_test_step(dev):
    from src.engines import VadModel
    g = torch.Generator().manual_seed(2)
    for n in (1, 15, 16, 1000, 123457):
        d = (torch.rand(n, generator=g) > 0.4).to(torch.uint8)
        y = (torch.rand(n, generator=g) > 0.6).to(torch.uint8)
        tp, fp, tn, fn = torch.ops.b200vad.stat_scores(d.to(dev), y.to(dev)).tolist()
        assert (tp, fp, tn, fn) == (int((d & y).sum()), int((d & (1 - y)).sum()), int(((1 - d) & (1 - y)).sum()), int(((1 - d) & y).sum()))
    feats = torch.randn(3, 200, 80, generator=g) * 3 - 5
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    y = (torch.rand(3, 200, generator=g) > 0.5).float()
    with torch.no_grad():
        r = m.test_step({"inputs": feats.to(dev), "is_voice": y.to(dev)})
        d = m.predict_step({"inputs": feats.to(dev)}).squeeze(-1).cpu()
    tp, fp, tn, fn = r["stat_scores"]
    assert (tp, fp, tn, fn) == (int(((d == 1) & (y == 1)).sum()), int(((d == 1) & (y == 0)).sum()),
                                int(((d == 0) & (y == 0)).sum()), int(((d == 0) & (y == 1)).sum()))
    assert r["test_false_alarm"] == fp / 600 and r["test_missed_detection"] == fn / 600


def test_detection_error_matches_reference_loop(dev):
    import oracle
    import b200vad
    from src.scripts.predict import get_binary_tensor, get_false_alarm, get_missed_detection, score_predictions
    rng = np.random.default_rng(4)
    fs = 0.01
    durations, gts, preds = [], [], []
    for r in range(9):
        dur = float(rng.uniform(0.5, 400.0)) if r else 3600.0
        def rand_ivs(k):
            pts = np.sort(rng.uniform(0, dur * 1.02, size=2 * k)).round(2)        # some intervals run past the end
            return [(float(pts[2 * i]), float(pts[2 * i + 1])) for i in range(k)]
        durations.append(dur); gts.append(rand_ivs(int(rng.integers(0, 40)))); preds.append(rand_ivs(int(rng.integers(0, 40))))
    # the reference's accumulation (predict.py:500-509, 590-600) on the oracle's CPU tensors
    fa_avg = md_avg = der_avg = 0
    for dur, g, p in zip(durations, gts, preds):
        gt_t, pr_t = oracle.get_binary_tensor(g, dur, fs), oracle.get_binary_tensor(p, dur, fs)
        fa, md = oracle.get_false_alarm(gt_t, pr_t), oracle.get_missed_detection(gt_t, pr_t)
        fa_avg += fa; md_avg += md; der_avg += fa + md
        # drop-in per-recording functions
        gt_d, pr_d = get_binary_tensor(g, dur, fs), get_binary_tensor(p, dur, fs)
        assert torch.equal(gt_d.cpu(), gt_t)
        assert torch.equal(get_false_alarm(gt_d, pr_d), fa) and torch.equal(get_missed_detection(gt_d, pr_d), md)
    n = len(durations)
    der, fa, md = score_predictions(gts, preds, durations, fs)
    assert torch.equal(fa, fa_avg / n) and torch.equal(md, md_avg / n) and torch.equal(der, der_avg / n)


# ---------------------------------------------------------------- corpus sharding (BASELINE config 4)
def test_corpus_sharding_is_invisible(dev):
    """Any sharding / batching of the synthetic corpus yields the same waveforms and the same global segment list."""
    import b200vad
    from b200vad import corpus
    U, N = 300, 32000
    a = corpus.synth_corpus(0, U, N, seed=7, device=dev)
    b = torch.cat([corpus.synth_corpus(0, 100, N, 7, dev).clone(), corpus.synth_corpus(100, 200, N, 7, dev).clone()])
    assert torch.equal(a, b) and torch.isfinite(a).all() and 0.005 < float(a.std()) < 0.2
    assert not torch.equal(a[0], a[1]) and not torch.equal(a, corpus.synth_corpus(0, U, N, seed=8, device=dev))
    feats = torch.ops.b200vad.fbank(a[:16], None).cpu()
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    whole, frames = corpus.run_corpus(blob, U, N, batch_rows=128, seed=7)
    assert frames == U * 200 and whole.shape[0] > 0
    for world in (2, 3, 8):
        parts = [corpus.run_corpus(blob, U, N, rank=r, world=world, batch_rows=64, seed=7, gather=False)[0] for r in range(world)]
        assert torch.equal(torch.cat(parts), whole), world
    ids = whole[:, 0].cpu()
    assert bool((ids[1:] >= ids[:-1]).all()) and int(ids.max()) < U
