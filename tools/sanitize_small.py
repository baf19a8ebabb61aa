"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the path once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad
from src.engines import VadModel
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
blob = b200vad.pack_model(m.model.state_dict(), dev, 80, 4)
wav = 0.1 * torch.randn(70, 16000, device=dev)           # 70 rows: one full + one partial 64-sequence block, T = 100
prob, dec, seg, counts = torch.ops.b200vad.vad_pipeline(wav, None, blob, 4, 0.5, 49)
feats = torch.ops.b200vad.fbank(wav[:3], torch.tensor([16000, 9000, 500], device=dev))
p2 = torch.ops.b200vad.lstm_head(feats, blob, 4)
tp = torch.ops.b200vad.stat_scores(dec, (prob > 0.4).to(torch.uint8))
lf = b200vad.LongFormVad(blob, 4, window=16000, hop=8000)(wav[:4].reshape(-1).contiguous())
sv = b200vad.StreamingVad(blob, 4, num_streams=3, window=8000, hop=160, use_graph=False)
for _ in range(3):
    sv.push(0.1 * torch.randn(3, 160))
sv.close()
pm = VadModel("PyanNet", {"encoding_dim": 60}).eval().to(dev)
with torch.no_grad():
    q = pm(wav[:2, None, :])
torch.cuda.synchronize()
print("sanitize run ok", float(prob.mean()), seg.shape[0], float(p2.mean()), tp.tolist(), len(lf["intervals"]), tuple(q.shape))
