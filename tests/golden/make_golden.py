"""Regenerates tests/golden/*.npz from the CPU oracle at fixed seeds.

The reference ships no golden vectors and its modules cannot be imported in this image, so
these fixtures come from the oracle restatement (same torch / scipy primitives).  They pin the oracle
against drift and give the GPU tests a box-independent target.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))

import oracle  # noqa: E402
import util  # noqa: E402


def main():
    torch.set_num_threads(1)
    wav = util.synth_wave(2, 16000, seed=5)
    feats = oracle.lhotse_fbank(wav)
    feats64 = oracle.lhotse_fbank(wav, dtype=torch.float64)
    np.savez_compressed(os.path.join(HERE, "fbank_small.npz"), wav=wav.numpy(), feats=feats.numpy(),
                        feats64=feats64.numpy().astype(np.float64))

    o = util.make_oracle("PyanNet2", {"encoding_dim": 80}, seed=42)
    with torch.no_grad():
        p = o(feats).squeeze(-1)
    os_ = util.make_oracle("PyanNet2", {"encoding_dim": 80}, seed=42, spread=True, feats=feats)
    with torch.no_grad():
        ps = os_(feats).squeeze(-1)
        ds = os_.predict_step({"inputs": feats}).squeeze(-1)
    segs = [oracle.rle_segments(ds[i].tolist(), 0.01) for i in range(2)]
    np.savez_compressed(os.path.join(HERE, "pyannet2_small.npz"), prob=p.numpy(), prob_spread=ps.numpy(),
                        dec_spread=ds.numpy().astype(np.uint8),
                        cls_w=os_.model.classifier.weight.detach().numpy(), cls_b=os_.model.classifier.bias.detach().numpy(),
                        seg0=np.array(segs[0], dtype=np.float64).reshape(-1, 2), seg1=np.array(segs[1], dtype=np.float64).reshape(-1, 2))

    o2 = util.make_oracle("PyanNet", {}, seed=42)
    with torch.no_grad():
        s = o2.model.sincnet(wav.unsqueeze(1))
        p2 = o2.model(wav.unsqueeze(1)).squeeze(-1)
        filt = o2.model.sincnet.conv1d[0].filterbank.filters()
    np.savez_compressed(os.path.join(HERE, "pyannet_small.npz"), sincnet=s.numpy(), prob=p2.numpy(), filters=filt.numpy())

    g = torch.Generator().manual_seed(9)
    prob = torch.sigmoid(torch.cumsum(torch.randn(4, 700, generator=g), 1) * 0.3)
    prob[0, 100] = 0.5
    m49 = oracle.median_filter(prob.clone(), window=0.01)
    m25 = oracle.median_filter(prob.clone(), window=0.02)
    frames = [np.array(oracle.postproc.rle_frames(m49[i].numpy()), dtype=np.int32).reshape(-1, 2) for i in range(4)]
    np.savez_compressed(os.path.join(HERE, "postproc.npz"), prob=prob.numpy(), med49=m49.numpy().astype(np.uint8),
                        med25=m25.numpy().astype(np.uint8), **{f"frames{i}": f for i, f in enumerate(frames)})
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
