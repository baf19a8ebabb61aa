"""Timings of the BASELINE configs that are not the bench line (parity for them lives in tests/test_gpu_modes.py):
  config 1: one 60 s clip as 12 reference windows through the drop-in modules (fbank -> VadModel.predict_step -> segments)
  config 3: 1 h of long-form audio, reference semantics (720 windows, hop = window) and overlapped (hop 2.5 s, 1439 windows)
  config 5: streaming, 256 streams x 5 s ring x 10 ms hop: per-push device time and host wall time, p50 / p99
Prints one JSON object; CUDA events around every call, host wall clock beside them."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad
from src.engines import VadModel
from src.features import Fbank, FbankConfig
from src.scripts.predict import get_segments

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = VadModel("PyanNet2", {"encoding_dim": 80}).eval().to(dev)
blob = b200vad.pack_model(model.model.state_dict(), dev, 80, 4)
out = {"gpu": torch.cuda.get_device_name(0)}


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev, wall = [], []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        ev.append(e0.elapsed_time(e1))
    ev.sort(); wall.sort()
    return {"device_ms_p50": ev[len(ev) // 2], "device_ms_min": ev[0], "wall_ms_p50": wall[len(wall) // 2]}


# ---- config 1
clip = b200vad.synth.meeting_batch(1, 960000, seed=42)[0].to(dev)
fb = Fbank(FbankConfig(device="cuda"))
def config1():
    rows = clip.view(12, 80000)
    with torch.no_grad():
        dec = model.predict_step({"inputs": fb.extract_batch(rows, 16000)}, 0)
    return get_segments(dec, [60.0], 0.01)
r = timed(config1, 20)
r["x_real_time"] = 60.0 / (r["wall_ms_p50"] / 1e3)
out["config1_60s_clip"] = r

# ---- config 3
N = 3600 * 16000
g = torch.Generator(device=dev).manual_seed(1)
wav = 0.05 * torch.randn(N, device=dev, generator=g)
for name, hop in (("reference_semantics_720_windows", None), ("overlap_hop_2.5s_1439_windows", 40000)):
    lf = b200vad.LongFormVad(blob, 4, hop=hop)
    r = timed(lambda: lf(wav), 5)
    r["x_real_time"] = 3600.0 / (r["wall_ms_p50"] / 1e3)
    out.setdefault("config3_one_hour", {})[name] = r
del wav

# ---- config 5
sv = b200vad.StreamingVad(blob, 4, num_streams=256, window=80000, hop=160)
gen = torch.Generator().manual_seed(0)
chunks = [(0.1 * torch.randn(256, 160, generator=gen)).pin_memory() for _ in range(64)]
dms, wms = [], []
for i in range(1100):
    t0 = time.perf_counter()
    prob, dec, ms = sv.push(chunks[i % 64])
    w = (time.perf_counter() - t0) * 1e3
    if i >= 100:
        dms.append(ms); wms.append(w)
sv.close()
dms.sort(); wms.sort()
q = lambda a, p: a[min(len(a) - 1, int(p * len(a)))]
out["config5_streaming_256x5s_hop10ms"] = {
    "pushes": len(dms), "device_ms_p50": q(dms, 0.5), "device_ms_p99": q(dms, 0.99), "device_ms_max": dms[-1],
    "wall_ms_p50": q(wms, 0.5), "wall_ms_p99": q(wms, 0.99), "wall_ms_max": wms[-1],
    "note": "host chunk (256 x 160 f32, pinned) -> ring append -> fbank + PyanNet2 + median over the 5 s window (CUDA graph) -> newest frame's "
            "probability / decision back on the host; real time needs <= 10 ms per push",
}
print(json.dumps(out))
