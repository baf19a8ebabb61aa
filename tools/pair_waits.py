"""Where the warps of the CTA-pair LSTM layer kernel (csrc/lstm_pair.cu) wait: per (CTA of cluster 0, warp, wait site) share of the
kernel's cycles spent in mbarrier waits that were not already satisfied, and the number of such waits (opt bit 64, probe build).

    python tools/pair_waits.py [B T] [extra opt bits, default 2]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "universal-voice-activity-detection_b200")):
    sys.path.insert(0, p)

import b200vad  # noqa: E402
import oracle  # noqa: E402

ROLES = {"pw": {6: "acc_ready"}, "prod": {1: "x_empty"}, "mmah": {2: "x_done", 4: "h_ready"},
         "mmax": {2: "acc_free", 3: "x_full", 5: "drain"}, "send": {1: "slice", 7: "h_free"}}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    extra = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dev = torch.device("cuda:0")
    lib = b200vad.lib()
    torch.manual_seed(42)
    m = oracle.VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    blob = b200vad.pack_model(m.model.state_dict(), dev, 80, 4)
    x = (torch.randn(B, T, 80, device=dev) * 3 - 5)
    lib.b200vad_set_lstm_fused(2)
    lib.b200vad_set_lstm_pair_opt(64 | extra)
    torch.ops.b200vad.lstm_head(x, blob, 4)
    torch.cuda.synchronize()
    n = 8 * 25 * 16
    buf = (C.c_longlong * n)()
    lib.b200vad_lstm_fused_read_debug(buf, n)
    print(f"--- wait sites of the last layer launch, opt {64 | extra} (share of the kernel's cycles in non-immediate waits; n = their count)")
    for cta in range(4):
        for warp, role in ((0, "pw"), (5, "pw"), (10, "pw"), (15, "pw"), (16, "prod"), (17, "mmah"), (18, "mmah"), (19, "mmax"), (20, "mmax"),
                           (21, "send"), (22, "send"), (24, "send")):
            base = (cta * 25 + warp) * 16
            tot = buf[base]
            names = ROLES[role]
            parts = [f"{names.get(t, t)} {buf[base + 2 * t] / max(tot, 1) * 100:5.1f}% (n={buf[base + 2 * t + 1]})" for t in range(1, 8) if buf[base + 2 * t + 1]]
            print(f"cta {cta} warp {warp:2d} {role:5s}: total {tot / 1e6:8.3f} Mcyc  " + "  ".join(parts), flush=True)
    lib.b200vad_set_lstm_pair_opt(3)
    lib.b200vad_set_lstm_fused(1)


if __name__ == "__main__":
    main()
