"""BASELINE config 4: a synthetic corpus sharded across the GPUs of one box (torchrun), NCCL gather of the segment
lists at the end.  Prints one JSON line on rank 0.  Usage: torchrun --nproc-per-node N tools/corpus_run.py --hours 1000"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "universal-voice-activity-detection_b200"))
import torch
import torch.distributed as dist
import b200vad
from b200vad import corpus
from src.engines import VadModel

ap = argparse.ArgumentParser()
ap.add_argument("--hours", type=float, default=1000.0)
ap.add_argument("--seconds", type=float, default=8.0)
ap.add_argument("--batch-rows", type=int, default=4096)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
model = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
blob = b200vad.pack_model(model.model.state_dict(), dev, 80, 4)
N = int(a.seconds * 16000)
U = int(round(a.hours * 3600 / a.seconds))
corpus.run_corpus(blob, min(U, a.batch_rows * world), N, rank, world, a.batch_rows)      # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
seg, frames = corpus.run_corpus(blob, U, N, rank, world, a.batch_rows)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": f"{a.hours:g} h synthetic corpus = {U} x {a.seconds:g} s utterances, generated on device, sharded x{world}",
                      "n_gpus": world, "seconds": t.item(), "audio_hours_per_s": a.hours / t.item(), "segments_gathered": int(seg.shape[0]),
                      "frames_rank0": frames, "includes": "waveform synthesis + fbank + PyanNet2 + median + segments + NCCL gather"}))
if world > 1:
    dist.destroy_process_group()
