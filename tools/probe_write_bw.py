"""Write-only HBM ceiling on this GPU: torch fill / cudaMemset of a 1.68 GB buffer (the size of
the cfg-2 paste output).  Context for the paste kernel's roofline fraction: MEASURED_PEAKS.json's
hbm_gbs is a COPY (read+write) figure."""
import torch

n = 32 * 100 * 512 * 1024
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("zero_", lambda: x.zero_()), ("fill_(1)", lambda: x.fill_(1)), ("copy_", lambda: y.copy_(x))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    bytes_moved = n * (2 if name == "copy_" else 1)
    print(f"{name:10s} {best:.4f} ms  {bytes_moved / best / 1e6:.0f} GB/s")
