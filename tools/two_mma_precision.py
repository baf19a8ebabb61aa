"""CPU study of the 2-MMA split-precision projection (gemm_xg2_kernel).  Layer outputs travel as two fp16 planes
y1 = fp16((1 - s) y), y2 = fp16(y - y1), s = 2^-6 (y1 + y2 = y to 2^-18); the projection of layers >= 1 is
xg = y1 . W_hi + y2 . W' with W' = fp16(W_hi + W_lo / s), one fp32 accumulator:
    y1.W_hi + y2.W' = y.((1 - s) W_hi + s W') + r.(W' - W_hi),  r = the rounding residual of y1 (2^-12 |y|),
and (1 - s) W_hi + s W' = W up to s * rounding(W') = 2^-18 |W|.  The recurrence (W_hh single fp16) and the head's
3-term product consume the same planes.  Compared against the 3-term product on hi / lo planes and the exact one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import oracle, util

torch.set_num_threads(8)
wav = util.synth_wave(8, 128000, seed=7)
feats = oracle.lhotse_fbank(wav).double()
S = 2.0 ** -6
def f16(x): return x.to(torch.float16).double()
def proj_exact(x, w): return x @ w.t()
def proj_3term(x, w):
    xh = f16(x); xl = f16(x - xh); wh = f16(w); wl = f16(w - wh)
    return xl @ wh.t() + xh @ wl.t() + xh @ wh.t()
def planes(x):
    x = x.float().double(); x1 = f16((1 - S) * x); return x1, f16(x - x1)
def proj_2mma(x, w):
    x1, x2 = planes(x); wh = f16(w); wl = f16(w - wh); wp = f16(wh + wl / S)
    return (x1 @ wh.t() + x2 @ wp.t()).float().double()
def proj_3term_planes(x, w):           # the 3-term kernel fed with (y1, y2) planes (the head)
    x1, x2 = planes(x); wh = f16(w); wl = f16(w - wh)
    return x2 @ wh.t() + x1 @ wl.t() + x1 @ wh.t()

def lstm_stack(sd, x, proj, proj0=None, h_planes=False):
    for l in range(4):
        outs = []
        for d, suf in enumerate(("", "_reverse")):
            wih = sd[f"model.lstm.weight_ih_l{l}{suf}"].double(); whh = f16(sd[f"model.lstm.weight_hh_l{l}{suf}"].double())
            b = (sd[f"model.lstm.bias_ih_l{l}{suf}"] + sd[f"model.lstm.bias_hh_l{l}{suf}"]).double()
            xg = (proj0 if (l == 0 and proj0 is not None) else proj)(x, wih) + b
            B, T, _ = x.shape
            h = torch.zeros(B, 128, dtype=torch.float64); c = torch.zeros_like(h)
            ys = [None] * T
            for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
                hq = sum(planes(h)) if h_planes else h
                g = xg[:, t] + hq @ whh.t()
                i, f, gg, o = g.chunk(4, 1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                ys[t] = h
            outs.append(torch.stack(ys, 1))
        x = torch.cat(outs, 2)
    return x
def head(sd, y, proj=proj_exact):
    z = torch.nn.functional.leaky_relu(proj(y, sd["model.linear.0.weight"].double()) + sd["model.linear.0.bias"].double())
    z = torch.nn.functional.leaky_relu(z @ sd["model.linear.1.weight"].double().t() + sd["model.linear.1.bias"].double())
    return torch.sigmoid(z @ sd["model.classifier.weight"].double().t() + sd["model.classifier.bias"].double()).squeeze(-1)

for sigma in (0.5, 2.0):
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats.float())
    with torch.no_grad():
        o.model.classifier.weight.mul_(sigma / 0.5); o.model.classifier.bias.mul_(sigma / 0.5)
    sd = o.state_dict()
    with torch.no_grad():
        # reference: everything exact except W_hh fp16 (the kernels' dominant rounding), to isolate the projection
        base = head(sd, lstm_stack(sd, feats, proj_exact))
        for name, pr, p0, hp, hd in (("3-term projection, hi / lo planes", proj_3term, None, False, proj_exact),
                                     ("2-MMA projection, all layers", proj_2mma, None, False, proj_exact),
                                     ("2-MMA on layers 1-3, 3-term on layer 0", proj_2mma, proj_3term, False, proj_exact),
                                     ("  + recurrence and head on the (y1, y2) planes", proj_2mma, proj_3term, True, proj_3term_planes)):
            p = head(sd, lstm_stack(sd, feats, pr, p0, hp), hd)
            print(f"sigma {sigma}: {name:50s} max rel err of p vs exact projection {((p - base).abs() / base).max():.2e}")
