"""CPU study of the operand-rounding choices of the LSTM stack (DESIGN.md "precision").

The kernels keep activations to ~22 bits (fp16 hi + lo planes) and accumulate in fp32, so their error against
the fp32 reference is governed by which WEIGHT matrices are rounded to a single fp16 plane.  This script rounds
the chosen matrices of a float64 copy of the reference-layout model and reports the max relative error of the
speech probability against the unrounded float64 model, for two logit spreads.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import copy
import torch
import oracle, util

torch.set_num_threads(8)
wav = util.synth_wave(16, 128000, seed=7)
feats = oracle.lhotse_fbank(wav)

def f16(w):
    return w.to(torch.float16).to(w.dtype)

def clone(o):
    m = oracle.VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    m.load_state_dict(o.state_dict())
    return m

def run(m, x):
    with torch.no_grad():
        return m(x).squeeze(-1)

for sigma in (0.5, 2.0):
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats)
    with torch.no_grad():
        o.model.classifier.weight.mul_(sigma / 0.5); o.model.classifier.bias.mul_(sigma / 0.5)
    ref = run(clone(o).double(), feats.double())
    p32 = run(o, feats)
    print(f"sigma={sigma}: p range [{ref.min():.3f}, {ref.max():.3f}]  fp32 torch vs fp64: {((p32.double()-ref).abs()/ref).max():.2e}")
    variants = {
        "W_hh f16 (current kernels)": lambda n: "weight_hh" in n,
        "W_hh f16 + W_ih f16 on layers 1-3": lambda n: "weight_hh" in n or ("weight_ih" in n and "_l0" not in n),
        "W_hh f16 + W_ih f16 on all layers": lambda n: "weight_hh" in n or "weight_ih" in n,
        "W_ih f16 on layers 1-3 only": lambda n: ("weight_ih" in n and "_l0" not in n),
        "W_hh + W_ih(1-3) + head f16": lambda n: "weight_hh" in n or ("weight_ih" in n and "_l0" not in n) or n.startswith("model.linear"),
    }
    for name, pick in variants.items():
        m = clone(o).double()
        with torch.no_grad():
            for n, p in m.named_parameters():
                if pick(n):
                    p.copy_(f16(p))
        p = run(m, feats.double())
        rel = ((p - ref).abs() / ref).max().item()
        logit = (torch.logit(p) - torch.logit(ref)).abs().max().item()
        print(f"   {name:42s} max rel err p {rel:.2e}   max |dlogit| {logit:.2e}")
