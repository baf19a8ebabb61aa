"""GPU parity at TRAINED-MODEL logit spreads: the random-init PyanNet2 / PyanNet of the named configs has
logits within 1e-3 of a constant, where every rounding choice looks exact.  Here the classifier is rescaled (a parameter
change, not a code change) so that the logits have standard deviation 2 and 4 -- p from ~1e-4 to ~1 - 1e-4 -- and the CUDA
path must still hold north_star's 1e-3 relative tolerance on p.  Decisions are compared outside the near-threshold band,
which is counted and printed.  Run on the B200 box: pytest -m gpu."""
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


def _run(dev, name, cfg, x, sigma, window):
    import oracle
    from src.engines import VadModel
    o = util.make_oracle(name, cfg, spread=True, feats=x, sigma=sigma)
    with torch.no_grad():
        ref_p = o(x.unsqueeze(1)) if name == "PyanNet" else o(x)
        ref_d = o.predict_step({"inputs": x})
    m = VadModel(name, dict(cfg)).eval()
    m.load_state_dict(o.state_dict())
    m = m.to(dev)
    with torch.no_grad():
        p = (m(x.to(dev).unsqueeze(1)) if name == "PyanNet" else m(x.to(dev))).cpu()
        d = m.predict_step({"inputs": x.to(dev)}).cpu()
    err = util.prob_err(p, ref_p)
    near = ((ref_p - 0.5).abs() <= util.NEAR_THR).squeeze(-1)
    diff = (d != ref_d).squeeze(-1)
    print(f"{name} {cfg} sigma={sigma}: rel err of p {err:.2e}; p in [{ref_p.min().item():.2e}, {ref_p.max().item():.6f}]; "
          f"near-threshold frames {int(near.sum())} of {near.numel()}; decisions differing {int(diff.sum())}")
    assert err <= util.PROB_RTOL, err
    # decisions from OUR probabilities through the oracle's median filter are ours bit-exactly ...
    assert torch.equal(oracle.median_filter(p.squeeze(-1), window=window), d.squeeze(-1))
    # ... and equal the oracle's own, except where a near-threshold frame sits inside the median window
    k = 49 if window == 0.01 else 25
    for b, t in zip(*torch.nonzero(diff, as_tuple=True)):
        assert near[b, max(0, int(t) - k // 2): int(t) + k // 2 + 1].any(), (int(b), int(t))
    return err


@pytest.mark.parametrize("sigma", [2.0, 4.0])
def test_pyannet2_fbank_width_trained_spread(dev, sigma):
    import oracle
    wav = util.synth_wave(12, 80000, seed=11)
    feats = oracle.lhotse_fbank(wav)
    _run(dev, "PyanNet2", {"encoding_dim": 80}, feats, sigma, 0.01)


@pytest.mark.parametrize("sigma", [2.0, 4.0])
def test_pyannet2_ssl_width_trained_spread(dev, sigma):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 150, 768, generator=g)
    _run(dev, "PyanNet2", {"encoding_dim": 768}, x, sigma, 0.02)


@pytest.mark.parametrize("sigma", [2.0, 4.0])
def test_pyannet_trained_spread(dev, sigma):
    wav = util.synth_wave(4, 48000, seed=13)
    _run(dev, "PyanNet", {}, wav, sigma, 0.01)


def test_legacy_path_reported(dev):
    """The round-1 path (projection GEMM -> xg -> recurrence with a single fp16 plane of W_hh) is kept for cross-validation;
    its error at sigma = 2 is printed next to the fused kernel's (it exceeded 1e-3 in the CPU emulation, tools/precision_study.py)."""
    import b200vad
    import oracle
    wav = util.synth_wave(12, 80000, seed=11)
    feats = oracle.lhotse_fbank(wav)
    lib = b200vad.lib()
    try:
        lib.b200vad_set_lstm_fused(0)
        o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats, sigma=2.0)
        from src.engines import VadModel
        with torch.no_grad():
            ref_p = o(feats)
        m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
        m.load_state_dict(o.state_dict())
        m = m.to(dev)
        with torch.no_grad():
            legacy = util.prob_err(m(feats.to(dev)).cpu(), ref_p)
    finally:
        lib.b200vad_set_lstm_fused(1)
    with torch.no_grad():
        fused = util.prob_err(m(feats.to(dev)).cpu(), ref_p)
    print(f"sigma=2: legacy path rel err {legacy:.2e}, fused path {fused:.2e}")
    assert fused <= util.PROB_RTOL
