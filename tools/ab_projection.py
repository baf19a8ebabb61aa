"""A/B of the input projections (development aid): single-CTA 3-term gemm_ts_kernel<3> / 2-MMA gemm_xg2_kernel against the
CTA-pair kernel (3-term on layer 0, 2 or 3 terms on layers >= 1) on the BASELINE shape, alternating in one process so
that all see the same clocks.  Prints average ms per launch of
the recurrence (kind 0), the projections (kind 1, all four layers) and the head (kind 2), and the step time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import ctypes as C
import torch
import b200vad
from b200vad import _lib
from src.engines import VadModel

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
torch.manual_seed(42)
m = VadModel("PyanNet2", {"encoding_dim": 80}).eval().cuda()
feats = torch.randn(rows, T, 80, device="cuda") * 3 - 5
L = _lib.lib()
probs = {}
with torch.no_grad():
    for rep in range(2):
        for pair, terms in ((0, 3), (1, 3), (1, 2), (2, 3), (2, 2)):
            _lib.check(L.b200vad_set_projection_terms(terms), "set_projection_terms")
            _lib.check(L.b200vad_set_projection_kernel(pair), "set_projection_kernel")
            m(feats); torch.cuda.synchronize()
            L.b200vad_profile_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): p = m(feats)
            e1.record(); torch.cuda.synchronize()
            probs[(pair, terms)] = p
            out = []
            for k in (0, 1, 2):
                t, n = C.c_double(0), C.c_int(0)
                L.b200vad_profile_collect(k, C.byref(t), C.byref(n))
                out.append("kind %d: %d x %.3f ms" % (k, n.value, t.value / max(n.value, 1)))
            L.b200vad_profile_enable(0)
            print("kernel %d terms %d | %s | forward %.2f ms" % (pair, terms, " | ".join(out), e0.elapsed_time(e1) / 3), flush=True)
ref = probs[(0, 3)]
for k, v in probs.items():
    print("max rel diff of p, kernel %d terms %d vs kernel 0: %.2e" % (k[0], k[1], ((v - ref).abs() / ref).max().item()))
