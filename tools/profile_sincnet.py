"""Driver for ncu: the SincNet front-end (PyanNet rows a3 / a4) at a given batch, two passes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad  # noqa: F401
from src.engines import VadModel
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(42)
m = VadModel("PyanNet", {"encoding_dim": 60}).eval().cuda()
wav = 0.1 * torch.randn(rows, 1, 128000, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = m.model.sincnet(wav)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.mean()))
