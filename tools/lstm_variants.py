"""Timing of the recurrence under env-selected variants (development aid): prints avg ms per launch, T = 800."""
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
feats = torch.randn(4096, 800, 80, device="cuda") * 3 - 5
L = _lib.lib()
with torch.no_grad():
    m(feats); torch.cuda.synchronize()
    L.b200vad_profile_enable(1)
    for _ in range(3): m(feats)
    torch.cuda.synchronize()
    out = []
    for k in (0, 1, 2):
        t, n = C.c_double(0), C.c_int(0)
        L.b200vad_profile_collect(k, C.byref(t), C.byref(n))
        out.append("kind %%d: %%d x %%.3f ms" %% (k, n.value, t.value / max(n.value, 1)))
    print(os.environ.get("VARIANT", ""), " | ".join(out))
''' % ROOT
for spec in sys.argv[1:]:
    env = dict(os.environ, VARIANT=spec)
    for kv in spec.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            env[k] = v
    subprocess.run([sys.executable, "-c", code], env=env, timeout=300)
