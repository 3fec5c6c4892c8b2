"""Duplex PCIe rate as a function of copy size (pinned memory, two streams)."""
import time
import torch

total = 512 * 1024 * 1024
h = torch.empty(total, dtype=torch.uint8, pin_memory=True); d = torch.empty(total, dtype=torch.uint8, device="cuda")
h2 = torch.empty(total, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(total, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(chunk, h2d=True, d2h=True, reps=4):
    def once():
        for off in range(0, total, chunk):
            if h2d:
                with torch.cuda.stream(s1):
                    d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h2[off:off + chunk].copy_(d2[off:off + chunk], non_blocking=True)
    once(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    return total / ((time.perf_counter() - t0) / reps) / 1e9


for chunk in (4 << 20, 16 << 20, 32 << 20, 128 << 20, 512 << 20):
    print(f"chunk {chunk >> 20:4d} MiB: H2D alone {run(chunk, True, False):5.1f}  D2H alone {run(chunk, False, True):5.1f}  "
          f"both (each) {run(chunk):5.1f} GB/s")
