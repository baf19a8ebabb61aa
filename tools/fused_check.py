"""GPU check of the fused LSTM layer kernel (csrc/lstm_fused.cu): PyanNet2 probabilities through torch.ops.b200vad.lstm_head
with the layer kernels (2 = CTA-pair fused, 1 = fused, 0 = projection + recurrence) against a float64 torch reference on the same device, for several (B, T, D, layers, logit spread).

    python tools/fused_check.py [quick]

Prints one line per case: max relative error of p for both paths, and the fused / legacy difference."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "universal-voice-activity-detection_b200")):
    sys.path.insert(0, p)

import b200vad  # noqa: E402
import oracle  # noqa: E402


def build(D, L, sigma, x):
    torch.manual_seed(42)
    m = oracle.VadModel("PyanNet2", {"encoding_dim": D, "lstm": {"hidden_size": 128, "num_layers": L, "bidirectional": True,
                                                                "monolithic": True, "dropout": 0.0}}).eval()
    net = m.model.double()
    with torch.no_grad():
        if sigma > 0:
            y, _ = net.lstm(x.double())
            for lin in net.linear:
                y = torch.nn.functional.leaky_relu(lin(y))
            z = net.classifier(y)
            scale = sigma / z.std().clamp_min(1e-9)
            net.classifier.bias.copy_((net.classifier.bias - z.mean()) * scale)
            net.classifier.weight.mul_(scale)
        ref = net(x.double())
    return m, ref


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    dev = torch.device("cuda:0")
    lib = b200vad.lib()
    print("fused clusters resident:", lib.b200vad_lstm_fused_clusters(), flush=True)
    cases = [(16, 8, 80, 1, 0.0), (16, 50, 80, 1, 0.0), (3, 100, 80, 2, 0.5), (33, 57, 80, 4, 0.5), (130, 40, 60, 4, 2.0),
             (64, 200, 80, 4, 2.0), (64, 200, 80, 4, 4.0), (300, 30, 256, 2, 2.0)]
    if not quick:
        cases += [(1024, 100, 80, 4, 2.0), (4096, 50, 80, 4, 4.0)]
    for (B, T, D, L, sigma) in cases:
        g = torch.Generator().manual_seed(B * 1000 + T)
        x = (torch.randn(B, T, D, generator=g) * 3 - 5) if D == 80 else torch.randn(B, T, D, generator=g)
        m, ref = build(D, L, sigma, x)
        blob = b200vad.pack_model({k: v.float() for k, v in m.model.state_dict().items()}, dev, D, L)
        xd = x.to(dev)
        out = {}
        for fused in (2, 1, 0):
            lib.b200vad_set_lstm_fused(fused)
            torch.cuda.synchronize()
            t0 = time.time()
            p = torch.ops.b200vad.lstm_head(xd, blob, L)
            torch.cuda.synchronize()
            out[fused] = (p.cpu().double(), time.time() - t0)
        lib.b200vad_set_lstm_fused(1)
        r = ref.squeeze(-1)
        e1 = ((out[1][0].reshape(r.shape) - r).abs() / r.abs()).max().item()
        e0 = ((out[0][0].reshape(r.shape) - r).abs() / r.abs()).max().item()
        e2 = ((out[2][0].reshape(r.shape) - r).abs() / r.abs()).max().item()
        d = (out[1][0] - out[0][0]).abs().max().item()
        d2 = (out[2][0] - out[1][0]).abs().max().item()
        print(f"B={B} T={T} D={D} L={L} sigma={sigma}: pair err {e2:.3e} ({out[2][1]*1e3:.1f} ms)  fused err {e1:.3e} ({out[1][1]*1e3:.1f} ms)  "
              f"legacy err {e0:.3e} ({out[0][1]*1e3:.1f} ms)  |fused-legacy| {d:.3e}  |pair-fused| {d2:.3e}  "
              f"p in [{r.min().item():.4f}, {r.max().item():.4f}]", flush=True)


if __name__ == "__main__":
    main()
