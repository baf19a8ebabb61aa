"""Minimal driver for ncu captures of the fused LSTM layer kernel: three PyanNet2 forwards on (B, T, 80) inputs.
    python tools/fused_prof.py [B T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "universal-voice-activity-detection_b200")):
    sys.path.insert(0, p)

import b200vad  # noqa: E402
import oracle  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = oracle.VadModel("PyanNet2", {"encoding_dim": 80}).eval()
blob = b200vad.pack_model(m.model.state_dict(), dev, 80, 4)
x = torch.randn(B, T, 80, device=dev) * 3 - 5
for _ in range(3):
    p = torch.ops.b200vad.lstm_head(x, blob, 4)
torch.cuda.synchronize()
print("ok", float(p.mean()))
