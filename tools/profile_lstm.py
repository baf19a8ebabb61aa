"""Small driver for ncu: one model forward (fbank features -> probabilities) at a given batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad
from src.engines import VadModel
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
L = b200vad.lib()
if len(sys.argv) > 3: L.b200vad_set_projection_kernel(int(sys.argv[3]))
if len(sys.argv) > 4: L.b200vad_set_projection_terms(int(sys.argv[4]))
torch.manual_seed(42)
m = VadModel("PyanNet2", {"encoding_dim": 80}).eval().cuda()
feats = torch.randn(rows, T, 80, device="cuda") * 3 - 5
with torch.no_grad():
    for _ in range(2):
        p = m(feats)
torch.cuda.synchronize()
print("ok", float(p.mean()))
