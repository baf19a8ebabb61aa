"""Per-stage timing of the PyanNet (SincNet front-end) path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch, ctypes as C
import b200vad
from b200vad import _lib
from src.engines import VadModel
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = VadModel("PyanNet", {"encoding_dim": 60}).eval().to(dev)
wav = 0.1 * torch.randn(rows, 1, 128000, device=dev)
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
with torch.no_grad():
    t_all = timeit(lambda: m(wav))
    t_sinc = timeit(lambda: m.model.sincnet(wav))
print(f"PyanNet {rows} x 8 s: total {t_all:.1f} ms (SincNet front-end {t_sinc:.1f} ms) -> {rows*8/3600/(t_all/1e3):.1f} audio-h/s")
L = _lib.lib()
L.b200vad_profile_enable(1)
with torch.no_grad():
    m(wav)
torch.cuda.synchronize()
for k in (0, 1, 2, 3):
    t, n = C.c_double(0), C.c_int(0)
    L.b200vad_profile_collect(k, C.byref(t), C.byref(n))
    print("kind", k, "launches", n.value, "total ms %.2f" % t.value)
