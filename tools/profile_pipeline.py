"""Driver for ncu: the whole device pipeline on the BASELINE shape (4096 x 8 s), two passes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad
from src.engines import VadModel
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(42)
m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
blob = b200vad.pack_model(m.model.state_dict(), torch.device("cuda:0"), 80, 4)
wav = 0.1 * torch.randn(rows, 128000, device="cuda")
for _ in range(2):
    prob, dec, seg, counts = torch.ops.b200vad.vad_pipeline(wav, None, blob, 4, 0.5, 49)
torch.cuda.synchronize()
print("ok", float(prob.mean()), seg.shape[0])
