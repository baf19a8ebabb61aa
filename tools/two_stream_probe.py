"""Device-resident steps of the bench workload issued on ONE stream versus alternately on TWO streams (two batches in flight):
the fused LSTM layer kernel occupies 132 of the 148 SMs (33 clusters of 4), so the other kernels of the NEXT batch (fbank, head,
post-processing) can run on the 16 SMs it leaves idle and in its wave tails.

    python tools/two_stream_probe.py [rows] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "universal-voice-activity-detection_b200")):
    sys.path.insert(0, p)

import b200vad  # noqa: E402
from b200vad import synth  # noqa: E402
from src.engines import VadModel  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    model = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    blob = b200vad.pack_model(model.model.state_dict(), dev, 80, 4)
    wav = synth.noise_batch(rows, 128000, seed=1234, pin=True).to(dev)
    ref = None
    configs = ((1, None), (2, None), (2, (-1, 0)), (3, None))
    if len(sys.argv) > 3:                                   # e.g. "2": only the two-stream configuration (stress runs)
        configs = tuple((int(c), None) for c in sys.argv[3].split(","))
    for nstreams, prio in configs:
        if prio:
            streams = [torch.cuda.Stream(priority=p) for p in prio]
        else:
            streams = [torch.cuda.Stream() for _ in range(nstreams)]
        outs = [None] * nstreams

        def step(i):
            s = streams[i % nstreams]
            with torch.cuda.stream(s):
                outs[i % nstreams] = torch.ops.b200vad.vad_pipeline_padded(wav, None, blob, 4, 0.5, 49)

        for i in range(4):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            step(i)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        p = outs[0][0]
        if ref is None:
            ref = p.clone()
        print(f"{nstreams} stream(s) prio {prio}: {ms:7.3f} ms per step = {rows * 8 / 3600 / (ms / 1e3):7.1f} audio-h/s   "
              f"max |p - p(1 stream)| {(p - ref).abs().max().item():.1e}", flush=True)


if __name__ == "__main__":
    main()
