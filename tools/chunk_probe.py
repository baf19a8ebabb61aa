"""Device-resident step of the bench workload as ONE pipeline call over 4096 rows versus two calls over 2048 rows (one wave of 32
LSTM work items each instead of 64 items over 33 clusters) versus four over 1024.   python tools/chunk_probe.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "universal-voice-activity-detection_b200")):
    sys.path.insert(0, p)

import b200vad  # noqa: E402
from b200vad import synth  # noqa: E402
from src.engines import VadModel  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
blob = b200vad.pack_model(model.model.state_dict(), dev, 80, 4)
wav = synth.noise_batch(4096, 128000, seed=1234, pin=True).to(dev)
for rep in range(2):
    for chunks in (1, 2, 4, 1, 2):
        rows = 4096 // chunks

        def step():
            for c in range(chunks):
                torch.ops.b200vad.vad_pipeline_padded(wav[c * rows:(c + 1) * rows], None, blob, 4, 0.5, 49)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{chunks} call(s) of {rows} rows: {ms:7.3f} ms per step = {4096 * 8 / 3600 / (ms / 1e3):7.1f} audio-h/s", flush=True)
