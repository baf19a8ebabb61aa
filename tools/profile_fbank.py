"""Small driver for ncu: the fused fbank kernel on the BASELINE shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wav = 0.1 * torch.randn(rows, 128000, device="cuda")
for _ in range(3):
    f = torch.ops.b200vad.fbank(wav, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f = torch.ops.b200vad.fbank(wav, None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"fbank {rows} x 8 s: {ms:.3f} ms  {(rows*128000*4 + rows*800*320)/ms/1e6:.1f} GB/s algorithmic", float(f.mean()))
