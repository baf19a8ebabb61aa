"""GPU tests of the tcgen05 kernels in isolation: the split-precision GEMM against an fp64 matmul and
the tcgen05 model path against the warp-MMA path and the oracle."""
import ctypes as C

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


def _linear(dev, a, w, bias, use_lo):
    import b200vad
    L = b200vad.lib()
    M, K = a.shape
    N = w.shape[0]
    c = torch.empty((M, N), dtype=torch.float32, device=dev)
    ws = torch.empty(4 * (M * K + N * (K + 64)) + 4096, dtype=torch.uint8, device=dev)
    b200vad._lib.check(L.b200vad_linear_split_f32(a.data_ptr(), M, K, w.data_ptr(), N, bias.data_ptr(), int(use_lo), c.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "linear_split")
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("M,K,N,use_lo", [(128, 64, 256, 1), (1000, 80, 1024, 1), (4096, 256, 1024, 0), (333, 128, 128, 1),
                                          (70000, 256, 1024, 0), (20001, 256, 1024, 1), (777, 256, 128, 1)])
def test_gemm_tc_matches_fp64(dev, M, K, N, use_lo):
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g) * (5.0 if K == 80 else 0.3)
    w = (torch.rand(N, K, generator=g) - 0.5) * 0.17
    bias = torch.randn(N, generator=g) * 0.1
    ref = a.double() @ w.double().t() + bias.double()
    c = _linear(dev, a.to(dev), w.to(dev), bias.to(dev), use_lo).cpu().double()
    scale = (a.double().abs() @ w.double().abs().t()).clamp_min(1e-6)
    err = ((c - ref).abs() / scale).max().item()
    # 3-term split: ~2^-21 relative to sum |a||w|; 2-term (no W_lo): weight rounding 2^-12 remains
    assert err < (2e-6 if use_lo else 3e-4), err


@pytest.mark.parametrize("B,T", [(3, 50), (64, 100), (130, 333)])
def test_tc_model_matches_warp_mma_and_oracle(dev, B, T):
    import b200vad
    import oracle
    wav = util.synth_wave(min(B, 4), T * 160, seed=T)
    feats = oracle.lhotse_fbank(wav)
    feats = feats.repeat((B + feats.shape[0] - 1) // feats.shape[0], 1, 1)[:B]
    feats = feats + 0.01 * torch.randn(feats.shape, generator=torch.Generator().manual_seed(B))
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    with torch.no_grad():
        ref = o(feats).squeeze(-1)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    L = b200vad.lib()
    if L.b200vad_set_impl(1) != 0:
        L.b200vad_set_impl(2)
        pytest.skip("the warp-MMA cross-validation kernels are built with `make VALIDATE=1` only")
    try:
        b200vad._lib.check(L.b200vad_set_impl(1), "set_impl")
        p1 = torch.ops.b200vad.lstm_head(feats.to(dev), blob, 4).cpu()
        b200vad._lib.check(L.b200vad_set_impl(2), "set_impl")
        p2 = torch.ops.b200vad.lstm_head(feats.to(dev), blob, 4).cpu()
    finally:
        L.b200vad_set_impl(2)
    e1, e2 = util.prob_err(p1, ref), util.prob_err(p2, ref)
    print(f"B={B} T={T}: warp-MMA rel err {e1:.2e}, tcgen05 rel err {e2:.2e}")
    assert e2 <= util.PROB_RTOL, e2
    assert e1 <= util.PROB_RTOL, e1


@pytest.mark.parametrize("B,T", [(5, 60), (70, 90), (200, 41)])
def test_recurrence_tile_variants_agree(dev, B, T):
    """16 and 64 sequences per CTA (latency / throughput schedules of the recurrence) give the same probabilities to
    fp32 rounding of the accumulation order, and both match the oracle."""
    import b200vad
    g = torch.Generator().manual_seed(B * T)
    x = torch.randn(B, T, 80, generator=g) * 3 - 5
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=x)
    with torch.no_grad():
        ref = o(x).squeeze(-1)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    L = b200vad.lib()
    out = {}
    try:
        for nb in (16, 64):
            b200vad._lib.check(L.b200vad_set_lstm_tile(nb), "set_lstm_tile")
            out[nb] = torch.ops.b200vad.lstm_head(x.to(dev), blob, 4).cpu()
    finally:
        L.b200vad_set_lstm_tile(0)
    assert util.prob_err(out[16], ref) <= util.PROB_RTOL and util.prob_err(out[64], ref) <= util.PROB_RTOL
    assert util.prob_err(out[16], out[64]) <= 1e-5
    assert L.b200vad_set_lstm_tile(32) != 0


@pytest.mark.parametrize("B,T", [(70, 64), (7, 333), (129, 41)])
def test_projection_kernels_agree(dev, B, T):
    """Input projections: the general kernel (0), the single-CTA resident-weight kernel (1) and the CTA-pair kernel (2,
    default) give bit-identical results with three fp16 products per k-step; with two products (layers >= 1: scaled
    (y1, y2) planes and W' = fp16(W_hi + 2^6 W_lo), the default) the probabilities move by ~2e-5 relative, far inside the
    tolerance, and every variant matches the oracle."""
    import b200vad
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, 80, generator=g) * 3 - 5
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=x)
    with torch.no_grad():
        ref = o(x).squeeze(-1)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    L = b200vad.lib()
    out = {}
    try:
        for kern, terms in ((0, 3), (1, 3), (2, 3), (1, 2), (2, 2)):
            b200vad._lib.check(L.b200vad_set_projection_kernel(kern), "set_projection_kernel")
            b200vad._lib.check(L.b200vad_set_projection_terms(terms), "set_projection_terms")
            out[(kern, terms)] = torch.ops.b200vad.lstm_head(x.to(dev), blob, 4).cpu()
    finally:
        L.b200vad_set_projection_kernel(2)
        L.b200vad_set_projection_terms(2)
    for k, v in out.items():
        e = util.prob_err(v, ref)
        print(f"B={B} T={T}: kernel {k[0]} terms {k[1]}: err vs oracle {e:.2e}")
        assert e <= util.PROB_RTOL
    assert torch.equal(out[(1, 3)], out[(0, 3)]) and torch.equal(out[(2, 3)], out[(0, 3)])
    assert torch.equal(out[(2, 2)], out[(1, 2)])
    assert util.prob_err(out[(2, 2)], out[(0, 3)]) <= 1e-4
    assert L.b200vad_set_projection_kernel(3) != 0 and L.b200vad_set_projection_terms(4) != 0


@pytest.mark.parametrize("B,T", [(70, 64), (7, 333), (1, 2), (129, 41)])
def test_fused_head_is_bit_identical(dev, B, T):
    """The one-kernel head (hidden activations in shared memory) issues the same products in the same order as the two-launch
    head (hidden activations as fp16 planes in HBM): identical probabilities, and both match the oracle."""
    import b200vad
    g = torch.Generator().manual_seed(3 * B + T)
    x = torch.randn(B, T, 80, generator=g) * 3 - 5
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=x)
    with torch.no_grad():
        ref = o(x).squeeze(-1)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    L = b200vad.lib()
    out = {}
    try:
        for fused in (0, 1):
            b200vad._lib.check(L.b200vad_set_head_fused(fused), "set_head_fused")
            out[fused] = torch.ops.b200vad.lstm_head(x.to(dev), blob, 4).cpu()
    finally:
        L.b200vad_set_head_fused(1)
    assert torch.equal(out[0], out[1])
    assert util.prob_err(out[1], ref) <= util.PROB_RTOL


def test_random_shapes_cross_variants(dev):
    """Seeded sweep over awkward shapes: the default kernels (CTA-pair projections, one-kernel head) against the general
    kernels (single-CTA 3-product projections, two-launch head) -- bit-identical with three products, within 1e-4 with two --
    and the fused SincNet convolutions against the unfused ones' oracle tolerance."""
    import random
    import b200vad
    rnd = random.Random(20261018)
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=torch.randn(4, 50, 80) * 3 - 5)
    blob = b200vad.pack_model(o.model.state_dict(), dev, 80, 4)
    L = b200vad.lib()
    shapes = [(rnd.randint(1, 200), rnd.randint(1, 90)) for _ in range(10)] + [(1, 1), (64, 1), (65, 2), (31, 3), (33, 127)]
    try:
        for B, T in shapes:
            x = (torch.randn(B, T, 80, generator=torch.Generator().manual_seed(B * 1000 + T)) * 3 - 5).to(dev)
            outs = {}
            for name, kern, terms, head in (("general", 0, 3, 0), ("pair3", 2, 3, 1), ("default", 2, 2, 1)):
                L.b200vad_set_projection_kernel(kern); L.b200vad_set_projection_terms(terms); L.b200vad_set_head_fused(head)
                outs[name] = torch.ops.b200vad.lstm_head(x, blob, 4).cpu()
            assert torch.isfinite(outs["default"]).all(), (B, T)
            assert torch.equal(outs["general"], outs["pair3"]), (B, T)
            assert util.prob_err(outs["default"], outs["general"]) <= 1e-4, (B, T)
    finally:
        L.b200vad_set_projection_kernel(2); L.b200vad_set_projection_terms(2); L.b200vad_set_head_fused(1)
    # SincNet: odd lengths around tile / pooling-group boundaries of the fused kernels
    op = util.make_oracle("PyanNet", {})
    from src.engines import VadModel
    m = VadModel("PyanNet", {}).eval()
    m.load_state_dict(op.state_dict())
    m = m.to(dev)
    for N in [1540 + 270 * k + rnd.randint(0, 269) for k in (0, 1, 5, 20, 41, 63, 64, 127)]:
        wav = util.synth_wave(2, N, seed=N)
        with torch.no_grad():
            ref = op.model.sincnet(wav.unsqueeze(1))
            got = m.model.sincnet(wav.to(dev).unsqueeze(1)).cpu()
        assert got.shape == ref.shape, N
        assert util.feat_err(got, ref) <= util.FEAT_RTOL, (N, util.feat_err(got, ref))
