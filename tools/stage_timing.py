"""Per-stage CUDA-event timing of the hot path (development aid; bench.py is the contract)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch

import b200vad
from src.engines import VadModel


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=8.0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    N = int(a.seconds * 16000)
    torch.manual_seed(42)
    m = VadModel("PyanNet2", {"encoding_dim": 80}).eval().to(dev)
    wav = 0.1 * torch.randn(a.rows, N, device=dev)
    hours = a.rows * a.seconds / 3600.0
    with torch.no_grad():
        t_fb = timeit(lambda: torch.ops.b200vad.fbank(wav, None))
        feats = torch.ops.b200vad.fbank(wav, None)
        t_model = timeit(lambda: m(feats))
        prob = m(feats).squeeze(-1)
        t_med = timeit(lambda: torch.ops.b200vad.threshold_median(prob, 0.5, 49, False))
        dec = torch.ops.b200vad.threshold_median(prob, 0.5, 49, False)
        t_seg = timeit(lambda: torch.ops.b200vad.segments(dec, None, 2))
        blob = m.model._packed.blob
        t_pipe = timeit(lambda: torch.ops.b200vad.vad_pipeline(wav, None, blob, 4, 0.5, 49))
    T = feats.shape[1]
    print(f"rows={a.rows} N={N} T={T} audio_hours={hours:.3f}")
    print(f"fbank   {t_fb:9.3f} ms  {(4*a.rows*N + 320*a.rows*T)/t_fb/1e6:8.1f} GB/s algorithmic")
    print(f"model   {t_model:9.3f} ms  {2.884e6*a.rows*T/t_model/1e9:8.1f} TFLOP/s algorithmic")
    print(f"median  {t_med:9.3f} ms")
    print(f"segment {t_seg:9.3f} ms")
    print(f"pipeline{t_pipe:9.3f} ms  -> {hours/(t_pipe/1e3):8.2f} audio-hours/s")


if __name__ == "__main__":
    main()
