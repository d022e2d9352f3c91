"""PCIe ceiling for the end-to-end loop (developer tool): concurrent pinned H2D + D2H of the C2 step's byte counts."""
import time
import torch

n = 182 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{mode}: {10 * n / dt / 1e9:.1f} GB/s per direction")
