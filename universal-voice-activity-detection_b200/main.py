"""Entry point mirroring main.py:12-55 for the in-scope dispatch (task="run", function="predict"):
one 60 s synthetic 16 kHz clip (BASELINE config 1) cut into the reference's 5 s windows, through
fbank -> PyanNet2 -> threshold / median -> segments on the GPU."""

import faulthandler

faulthandler.enable()
import torch

from config.config import load_config


def main(config):
    if config.task != "run" or config.function not in ("predict", "predict_sincnet"):
        raise NotImplementedError(f"task={config.task!r} function={config.function!r} is outside the accelerated hot path")
    import b200vad
    from src.engines import VadModel
    from src.scripts.predict import get_segments, predict_vad

    torch.manual_seed(config.seed)
    model = VadModel(config.model_name, dict(config.model_dict)).eval()
    if config.load_checkpoint:
        ckpt = torch.load(config.checkpoint_path, map_location="cpu")
        model.load_state_dict(ckpt.get("state_dict", ckpt), strict=False)
    model = model.cuda()
    n = int(config.clip_seconds * 16000)
    win = int(config.window_seconds * 16000)
    clip = b200vad.synth.meeting_batch(1, n, seed=config.seed)[0]
    rows = clip[: (n // win) * win].view(-1, win).cuda()
    preds, _ = predict_vad(model, rows, frame_shift=config.frame_shift)
    sincnet = config.model_name == "PyanNet"
    segs = get_segments(preds, [config.clip_seconds], config.frame_shift, sincnet=sincnet)
    print(f"recording of {config.clip_seconds:.0f} s -> {len(segs[0])} speech segments: {segs[0][:8]}")
    return segs


if __name__ == "__main__":
    print("GPU:", torch.cuda.is_available())
    main(load_config())
