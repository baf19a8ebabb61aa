"""GPU parity: the CUDA path (through the C ABI / torch.library ops / drop-in modules) against the
CPU oracle on the same seeded inputs.  Run on the B200 box: pytest -m gpu."""
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


# ---------------------------------------------------------------- a1 fbank
@pytest.mark.parametrize("B,N", [(3, 16000), (2, 80000), (1, 960000), (4, 12345), (2, 400), (1, 161)])
def test_fbank_matches_oracle(dev, B, N):
    import oracle
    wav = util.synth_wave(B, N, seed=N)
    ref = oracle.lhotse_fbank(wav)
    ref64 = oracle.lhotse_fbank(wav, dtype=torch.float64)
    out = torch.ops.b200vad.fbank(wav.to(dev), None).cpu()
    assert out.shape == ref.shape
    # fp32 reference's own error against fp64 sets the scale; ours must be within tolerance of both
    assert util.feat_err(out, ref64) <= util.FEAT_RTOL, util.feat_err(out, ref64)
    assert util.feat_err(out, ref) <= 2 * util.FEAT_RTOL


def test_fbank_white_noise_and_lens(dev):
    import oracle
    torch.manual_seed(1)
    wav = 0.1 * torch.randn(5, 48000)
    lens = torch.tensor([48000, 47999, 32000, 16080, 1000], dtype=torch.int32)
    ref = oracle.lhotse_fbank(wav, lens=lens)
    out = torch.ops.b200vad.fbank(wav.to(dev), lens.to(dev)).cpu()
    assert util.feat_err(out, ref) <= util.FEAT_RTOL
    # rows are independent: a strided (non-contiguous-row) view gives the same result
    big = torch.zeros(5, 50000)
    big[:, :48000] = wav
    out2 = torch.ops.b200vad.fbank(big.to(dev)[:, :48000], lens.to(dev)).cpu()
    assert torch.equal(out, out2)


def test_fbank_linearity_full_size(dev):
    """Size-independent property at the BASELINE shape (8 s rows): scaling the waveform by 2 shifts every
    log-mel value by ln 4 (where above the eps floor), and rows do not interact."""
    wav = 0.05 * torch.randn(256, 128000, device=dev)
    a = torch.ops.b200vad.fbank(wav, None)
    b = torch.ops.b200vad.fbank(2 * wav, None)
    assert a.shape == (256, 800, 80)
    live = a > -15.0   # a few low mel bins contain no FFT bin and sit at log(eps) in both (as in Kaldi)
    assert live.float().mean().item() > 0.9
    assert (b - a - np.log(4.0))[live].abs().max().item() < 2e-4
    c = torch.ops.b200vad.fbank(wav[17:18].contiguous(), None)
    assert torch.equal(c[0], a[17])


# ---------------------------------------------------------------- a2 LSTM stack + head
@pytest.mark.parametrize("B,T", [(3, 100), (33, 57), (1, 500), (64, 800)])
def test_pyannet2_probabilities(dev, B, T):
    import oracle
    from src.engines import VadModel
    wav = util.synth_wave(min(B, 4), T * 160, seed=T)
    feats = oracle.lhotse_fbank(wav)
    feats = feats.repeat((B + feats.shape[0] - 1) // feats.shape[0], 1, 1)[:B]
    feats = feats + 0.01 * torch.randn(feats.shape, generator=torch.Generator().manual_seed(B))
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80})
    with torch.no_grad():
        ref = o(feats)
    m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    with torch.no_grad():
        out = m(feats.to(dev)).cpu()
    assert out.shape == ref.shape == (B, T, 1)
    assert util.prob_err(out, ref) <= util.PROB_RTOL, util.prob_err(out, ref)


def test_pyannet2_ssl_width(dev):
    """The reference's default configuration feeds 768-dim SSL features (config/config.py:16-18, 30-35,
    VadModel defaults vad_engine.py:30-35): the layer-0 projection takes the wide-input kernel path."""
    from src.engines import VadModel
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 120, 768, generator=g)
    o = util.make_oracle("PyanNet2", {"encoding_dim": 768}, spread=True, feats=x)
    with torch.no_grad():
        ref = o(x)
    m = VadModel("PyanNet2", {"encoding_dim": 768}).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    with torch.no_grad():
        out = m(x.to(dev)).cpu()
    assert util.prob_err(out, ref) <= util.PROB_RTOL, util.prob_err(out, ref)


def test_pyannet2_spread_head_decisions(dev):
    """Spread-head variant: probabilities span (0,1); decisions / segments must be bit-exact except
    frames within the stated tolerance of the threshold, which are counted separately."""
    import oracle
    from src.engines import VadModel
    from src.scripts.predict import get_segments
    wav = util.synth_wave(12, 80000, seed=7)
    feats = oracle.lhotse_fbank(wav)
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    with torch.no_grad():
        ref_p = o(feats)
        ref_d = o.predict_step({"inputs": feats})
    assert ref_p.min() < 0.4 and ref_p.max() > 0.6, (ref_p.min(), ref_p.max())
    m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    with torch.no_grad():
        p = m(feats.to(dev)).cpu()
        d = m.predict_step({"inputs": feats.to(dev)})
    assert d.dtype == torch.int64 and d.shape == (12, 500, 1) and d.is_cuda
    err = util.prob_err(p, ref_p)
    assert err <= util.PROB_RTOL, err
    near = ((ref_p - 0.5).abs() <= util.NEAR_THR).squeeze(-1)
    # decisions re-derived from OUR probabilities through the oracle's median filter must equal ours bit-exactly
    assert torch.equal(oracle.median_filter(p.squeeze(-1), window=0.01), d.cpu().squeeze(-1))
    if near.sum() == 0:
        assert torch.equal(d.cpu(), ref_d)
    segs = get_segments(d, None, 0.01)
    for i in range(12):
        want = oracle.merge_intervals_with_buffer(oracle.rle_segments(d[i, :, 0].cpu().tolist(), 0.01), 5.0, 0)
        assert [list(x) for x in segs[i]] == [list(x) for x in want]
    print(f"near-threshold frames: {int(near.sum())} of {near.numel()}; prob rel err {err:.2e}")


# ---------------------------------------------------------------- a3/a4 SincNet + PyanNet
# short inputs: a single pooled tile per layer / a single output frame (tile and pooling-group boundaries of the fused kernels)
@pytest.mark.parametrize("B,N", [(2, 80000), (3, 16000), (1, 128000), (2, 2000), (1, 1540), (5, 12345)])
def test_sincnet_and_pyannet(dev, B, N):
    from src.engines import VadModel
    wav = util.synth_wave(B, N, seed=N + 1)
    o = util.make_oracle("PyanNet", {})
    with torch.no_grad():
        ref_s = o.model.sincnet(wav.unsqueeze(1))
        ref_p = o.model(wav.unsqueeze(1))
    m = VadModel("PyanNet", {}).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    with torch.no_grad():
        s = m.model.sincnet(wav.to(dev).unsqueeze(1)).cpu()
        p = m.model(wav.to(dev).unsqueeze(1)).cpu()
    assert s.shape == ref_s.shape and p.shape == ref_p.shape
    # SincNet outputs are instance-normalised (unit scale): absolute tolerance 1e-3 * max(1, |ref|)
    assert util.feat_err(s, ref_s) <= util.FEAT_RTOL, util.feat_err(s, ref_s)
    assert util.prob_err(p, ref_p) <= util.PROB_RTOL, util.prob_err(p, ref_p)


# ---------------------------------------------------------------- a6/a7 threshold + median
@pytest.mark.parametrize("B,T,window", [(7, 500, 0.01), (3, 250, 0.02), (2, 4999, 0.01), (1, 10, 0.01), (5, 49, 0.01)])
def test_median_filter_bit_exact(dev, B, T, window):
    import oracle
    from src.utils.helper import median_filter
    g = torch.Generator().manual_seed(T)
    # smooth random walk so runs have realistic lengths, plus exact 0.5 and NaN entries
    x = torch.sigmoid(torch.cumsum(torch.randn(B, T, generator=g), 1) * 0.3)
    x[0, T // 2] = 0.5
    x[-1, T // 3] = float("nan")
    ref = oracle.median_filter(x.clone(), window=window)
    out = median_filter(x.to(dev), window=window)
    assert out.dtype == torch.int64 and out.is_cuda
    assert torch.equal(out.cpu(), ref)


# ---------------------------------------------------------------- a8/a9 segments
def test_segments_bit_exact(dev):
    import oracle
    from src.scripts.predict import get_segments
    g = torch.Generator().manual_seed(3)
    d = (torch.rand(40, 1300, generator=g) < 0.5).to(torch.uint8)
    d[1] = 1
    d[2] = 0
    d[3, :700] = 1
    d[3, 700:] = 0
    d[4, ::2] = 1
    d[4, 1::2] = 0
    # long runs crossing the 256-frame chunks of the kernel
    d[5] = 0
    d[5, 100:900] = 1
    d[5, 1290:] = 1
    got = get_segments(d.to(dev), None, 0.01)
    for i in range(d.shape[0]):
        want = oracle.merge_intervals_with_buffer(oracle.rle_segments(d[i].tolist(), 0.01), 13.0, 0)
        assert [list(x) for x in got[i]] == [list(x) for x in want], i
    # per-recording re-slicing of the flat stream (predict.py:447-458)
    durations = [4.99, 13.0, 0.5, 20.2, 7.77]
    got = get_segments(d.to(dev), durations, 0.01)
    streams = oracle.slice_recordings(d.reshape(-1), durations, 0.01)
    for i, s in enumerate(streams):
        want = oracle.merge_intervals_with_buffer(oracle.rle_segments(s.tolist(), 0.01), durations[i], 0)
        assert [list(x) for x in got[i]] == [list(x) for x in want], i


def test_segments_sincnet_time_base(dev):
    import oracle
    from src.scripts.predict import get_segments
    g = torch.Generator().manual_seed(5)
    d = (torch.sigmoid(torch.cumsum(torch.randn(6, 293, generator=g), 1)) > 0.5).to(torch.uint8)
    # without explicit durations a row's duration is what its frame count covers: 270-sample frames, 991-sample receptive field
    dur = ((293 - 1) * 270 + 991) / 16000.0
    got = get_segments(d.to(dev), None, 0.02, sincnet=True)
    for i in range(6):
        want = oracle.merge_intervals_with_buffer(oracle.rle_segments_sincnet(d[i].tolist(), dur), dur, 0)
        assert [list(x) for x in got[i]] == [list(x) for x in want], i
    # an explicit per-row duration (what predict_vad passes: samples / 16000)
    got = get_segments(d.to(dev), None, 0.02, sincnet=True, row_duration=5.0)
    for i in range(6):
        want = oracle.merge_intervals_with_buffer(oracle.rle_segments_sincnet(d[i].tolist(), 5.0), 5.0, 0)
        assert [list(x) for x in got[i]] == [list(x) for x in want], i


# ---------------------------------------------------------------- whole path, C ABI pipeline and host session
def test_pipeline_and_host_session(dev):
    import b200vad
    import oracle
    wav = util.synth_wave(10, 80000, seed=11)
    feats = oracle.lhotse_fbank(wav)
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    with torch.no_grad():
        ref_p = o(feats).squeeze(-1)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    prob, dec, seg, counts = torch.ops.b200vad.vad_pipeline(wav.to(dev), None, blob, 4, 0.5, 49)
    assert util.prob_err(prob.cpu(), ref_p) <= util.PROB_RTOL
    assert torch.equal(oracle.median_filter(prob.cpu(), window=0.01), dec.cpu().long())
    want = [(i, a, b) for i in range(10) for a, b in oracle.postproc.rle_frames(dec[i].cpu().numpy())]
    assert [tuple(r) for r in seg.cpu().tolist()] == want
    assert counts.cpu().tolist() == [sum(1 for w in want if w[0] == i) for i in range(10)]
    # host-buffer session with small chunks (exercises the double-buffered H2D pipeline)
    sess = b200vad.HostSession(blob, 4, 80000, chunk_rows=4)
    res = sess.run(wav.pin_memory(), 0.5, 49, want_dec=True, want_prob=True)
    assert torch.equal(res["dec"][:10], dec.cpu())
    assert torch.allclose(res["prob"][:10], prob.cpu(), rtol=0, atol=0)
    assert [tuple(r) for r in res["seg"].tolist()] == want
    sess.close()
    # asynchronous form: two batches in flight, results identical to the device pipeline
    sess = b200vad.HostSession(blob, 4, 80000, chunk_rows=6)
    halves = [wav[:6].contiguous().pin_memory(), wav[6:].contiguous().pin_memory()]
    outs = [sess.submit(i, halves[i], 0.5, 49, want_dec=True, want_prob=True) for i in range(2)]
    with pytest.raises(b200vad.B200VadError):
        sess.submit(0, halves[0])                 # slot 0 still holds an unwaited batch
    for rep in range(2):                          # second round re-uses the slots
        for i, (lo, hi) in enumerate(((0, 6), (6, 10))):
            r = sess.wait(i, outs[i])
            assert torch.equal(r["dec"], dec[lo:hi].cpu())
            assert torch.equal(r["prob"], prob[lo:hi].cpu())
            assert [(a + lo, b, c) for a, b, c in r["seg"].tolist()] == [w for w in want if lo <= w[0] < hi]
            if rep == 0:
                outs[i] = sess.submit(i, halves[i], 0.5, 49, want_dec=True, want_prob=True, out=outs[i])
    with pytest.raises(b200vad.B200VadError):
        sess.wait(0, outs[0])                     # nothing in flight
    sess.close()


def test_config1_main_clip(dev):
    """BASELINE config 1: random-init model (seed 42) on one 60 s clip cut into 12 x 5 s windows."""
    import oracle
    from config.config import load_config
    from src.engines import VadModel
    import b200vad
    cfg = load_config("fbank")
    clip = b200vad.synth.meeting_batch(1, 960000, seed=42)[0]
    rows = clip.view(12, 80000)
    o = util.make_oracle("PyanNet2", dict(cfg.model_dict), seed=cfg.seed)
    feats = oracle.lhotse_fbank(rows)
    with torch.no_grad():
        ref_p = o(feats)
        ref_d = o.predict_step({"inputs": feats})
    m = VadModel(cfg.model_name, dict(cfg.model_dict)).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    from src.features import Fbank, FbankConfig
    f = Fbank(FbankConfig(device="cuda")).extract_batch(rows.to(dev), 16000)
    with torch.no_grad():
        p = m(f).cpu()
        d = m.predict_step({"inputs": f}).cpu()
    assert util.prob_err(p, ref_p) <= util.PROB_RTOL
    near = int(((ref_p - 0.5).abs() <= util.NEAR_THR).sum())
    if near == 0:
        assert torch.equal(d, ref_d)
