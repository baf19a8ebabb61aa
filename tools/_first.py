import os, sys, torch, time, ctypes as C
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "universal-voice-activity-detection_b200")): sys.path.insert(0, p)
import b200vad, oracle
dev=torch.device("cuda:0")
torch.manual_seed(42)
lib = b200vad.lib()
D, B, T, L, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
if mode == "head2": lib.b200vad_set_head_fused(0)
if mode == "legacy": lib.b200vad_set_lstm_fused(0)
if mode.startswith("flags"): lib.b200vad_set_lstm_fused_debug(int(mode[5:]), -1)
rec = lambda: (lambda b: (lib.b200vad_lstm_fused_last_timeout(b), list(b))[1])((C.c_int*7)())
i = -1; t0 = time.time()
try:
    m = oracle.VadModel("PyanNet2", {"encoding_dim": D, "lstm": {"hidden_size": 128, "num_layers": L, "bidirectional": True, "monolithic": True, "dropout": 0.0}}).eval()
    blob = b200vad.pack_model(m.model.state_dict(), dev, D, L)
    x = torch.randn(B, T, D, device=dev)
    torch.cuda.synchronize()
    for i in range(16):
        t0 = time.time()
        p = torch.ops.b200vad.lstm_head(x, blob, L); torch.cuda.synchronize()
    print("ok", flush=True)
except Exception as e:
    print("FAILED iter", i, "after %.2f s" % (time.time() - t0), rec(), str(e)[:300], flush=True)
