"""CPU study: how many bits does the stored input projection xg need?  Manual float64 LSTM with xg quantised."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import oracle, util

torch.set_num_threads(8)
wav = util.synth_wave(8, 128000, seed=7)
feats = oracle.lhotse_fbank(wav).double()

def q_f16(x): return x.to(torch.float16).double()
def q_bf16(x): return x.to(torch.bfloat16).double()
def q_bits(bits):
    def f(x):
        m, e = torch.frexp(x)
        return torch.ldexp(torch.round(m * 2 ** bits) / 2 ** bits, e)
    return f
def q_f16_hi_lo8(x):
    hi = x.to(torch.float16).double()
    lo = x - hi
    # lo stored as e5m2-ish: 3 significant bits
    m, e = torch.frexp(lo)
    return hi + torch.ldexp(torch.round(m * 8) / 8, e)

def lstm_stack(sd, x, quant, whh_f16=True):
    for l in range(4):
        outs = []
        for d, suf in enumerate(("", "_reverse")):
            wih = sd[f"model.lstm.weight_ih_l{l}{suf}"].double(); whh = sd[f"model.lstm.weight_hh_l{l}{suf}"].double()
            if whh_f16: whh = whh.to(torch.float16).double()
            b = (sd[f"model.lstm.bias_ih_l{l}{suf}"] + sd[f"model.lstm.bias_hh_l{l}{suf}"]).double()
            xg = quant(x @ wih.t() + b)
            B, T, _ = x.shape
            h = torch.zeros(B, 128, dtype=torch.float64); c = torch.zeros_like(h)
            ys = [None] * T
            rng = range(T) if d == 0 else range(T - 1, -1, -1)
            for t in rng:
                g = xg[:, t] + h @ whh.t()
                i, f, gg, o = g.chunk(4, 1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                ys[t] = h
            outs.append(torch.stack(ys, 1))
        x = torch.cat(outs, 2)
    return x

def head(sd, y):
    z = torch.nn.functional.leaky_relu(y @ sd["model.linear.0.weight"].double().t() + sd["model.linear.0.bias"].double())
    z = torch.nn.functional.leaky_relu(z @ sd["model.linear.1.weight"].double().t() + sd["model.linear.1.bias"].double())
    return torch.sigmoid(z @ sd["model.classifier.weight"].double().t() + sd["model.classifier.bias"].double()).squeeze(-1)

for sigma in (0.5, 2.0):
    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, spread=True, feats=feats.float())
    with torch.no_grad():
        o.model.classifier.weight.mul_(sigma / 0.5); o.model.classifier.bias.mul_(sigma / 0.5)
    sd = o.state_dict()
    with torch.no_grad():
        ref = head(sd, lstm_stack(sd, feats, lambda x: x, whh_f16=False))
        base = head(sd, lstm_stack(sd, feats, lambda x: x))
        print(f"sigma {sigma}: W_hh f16 only: {((base-ref).abs()/ref).max():.2e}")
        for name, q in (("xg fp16", q_f16), ("xg bf16", q_bf16), ("xg 16-bit mantissa (3 bytes)", q_bits(16)),
                        ("xg 13-bit mantissa", q_bits(13)), ("xg fp16 + 3-bit lo (3 bytes)", q_f16_hi_lo8)):
            p = head(sd, lstm_stack(sd, feats, q))
            print(f"   + {name:32s} {((p-ref).abs()/ref).max():.2e}")
