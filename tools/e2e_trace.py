"""Diagnostic: (1) H2D bandwidth while the device pipeline runs on another stream; (2) host timestamps of the
submit / wait calls of the pipelined host session."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import b200vad
from b200vad import synth
from src.engines import VadModel

rows, N = 4096, 128000
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
blob = b200vad.pack_model(model.model.state_dict(), dev, 80, 4)
wav_host = synth.noise_batch(rows, N, seed=1, pin=True)
wav_dev = wav_host.to(dev)
dst = torch.empty_like(wav_dev)

def pipe():
    return torch.ops.b200vad.vad_pipeline(wav_dev, None, blob, 4, 0.5, 49)

for _ in range(2):
    pipe()
torch.cuda.synchronize()
side = torch.cuda.Stream()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
# alone
with torch.cuda.stream(side):
    ev[0].record(); dst.copy_(wav_host, non_blocking=True); ev[1].record()
torch.cuda.synchronize()
print(f"H2D alone: {ev[0].elapsed_time(ev[1]):.1f} ms")
# under load
ev[2].record()
for _ in range(3):
    pipe()
ev[3].record()
with torch.cuda.stream(side):
    ev[0].record(); dst.copy_(wav_host, non_blocking=True); ev[1].record()
torch.cuda.synchronize()
print(f"H2D under load: {ev[0].elapsed_time(ev[1]):.1f} ms; 3 pipeline passes with the copy running: {ev[2].elapsed_time(ev[3]):.1f} ms")
ev[2].record()
for _ in range(3):
    pipe()
ev[3].record()
torch.cuda.synchronize()
print(f"3 pipeline passes alone: {ev[2].elapsed_time(ev[3]):.1f} ms")

sess = b200vad.HostSession(blob, 4, N, chunk_rows=rows, device=0)
outs = [{}, {}]
def run(k, log):
    t0 = time.perf_counter()
    for i in range(k):
        outs[i & 1] = sess.submit(i & 1, wav_host, 0.5, 49, want_dec=True, out=outs[i & 1])
        log.append(("submit", i, (time.perf_counter() - t0) * 1e3))
        if i >= 1:
            sess.wait((i - 1) & 1, outs[(i - 1) & 1]); log.append(("wait", i - 1, (time.perf_counter() - t0) * 1e3))
            log.append(("times", i - 1, sess.slot_times((i - 1) & 1)))
    sess.wait((k - 1) & 1, outs[(k - 1) & 1]); log.append(("wait", k - 1, (time.perf_counter() - t0) * 1e3))
    log.append(("times", k - 1, sess.slot_times((k - 1) & 1)))
run(2, [])
log = []
run(5, log)
base = None
for r in log:
    if r[0] == "times":
        base = r[2][0] if base is None else base
        print("   slot timeline: h2d %.1f-%.1f  compute %.1f-%.1f  d2h end %.1f" % tuple(x - base for x in r[2]))
    else:
        print("%-7s %d  %8.1f ms" % r)
sess.close()
