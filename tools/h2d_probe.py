"""PCIe probe: pinned-host -> device copy bandwidth at the bench's transfer sizes (CUDA events)."""
import torch
for mb in (64, 512, 2048):
    h = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    e0.record()
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 3
    print(f"{mb} MiB: H2D {mb * 1.048576 / ms:.1f} GB/s  D2H {mb * 1.048576 / ms2:.1f} GB/s")
