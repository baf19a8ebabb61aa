"""Ablation timing of the tcgen05 recurrent kernel (development aid): B200VAD_LSTM_DEBUG flags
1 = no xg loads, 2 = no MMAs, 4 = no h_lo MMAs, 8 = no h smem writes."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "universal-voice-activity-detection_b200"))
import torch, ctypes as C
import b200vad
from b200vad import _lib
from src.engines import VadModel
torch.manual_seed(42)
m = VadModel("PyanNet2", {"encoding_dim": 80}).eval().cuda()
feats = torch.randn(4096, 400, 80, device="cuda") * 3 - 5
L = _lib.lib()
with torch.no_grad():
    m(feats); torch.cuda.synchronize()
    L.b200vad_profile_enable(1)
    for _ in range(2): m(feats)
    torch.cuda.synchronize()
    for k in (0, 1):
        t, n = C.c_double(0), C.c_int(0)
        L.b200vad_profile_collect(k, C.byref(t), C.byref(n))
        print("flags", os.environ.get("B200VAD_LSTM_DEBUG", "0"), "kind", k, "launches", n.value, "avg ms", t.value / max(n.value, 1))
''' % ROOT
for flags in sys.argv[1:]:
    env = dict(os.environ, B200VAD_LSTM_DEBUG=flags)
    subprocess.run([sys.executable, "-c", code], env=env, timeout=200)
