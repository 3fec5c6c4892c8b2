"""PCIe rate of every GPU of the box AT ONCE (pinned memory, one process, one thread and two streams per GPU):
does the host side of an 8-GPU box carry 8 links' worth of copies?  (round-1 e2e weak-scaling efficiency was
0.23 at N=8; this separates the fabric from the library's upload / download code.)

    python exp/bw_all.py [n_gpus]
"""
import sys
import threading
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
total = 512 * 1024 * 1024
chunk = 32 << 20
bufs = []
for g in range(n):
    with torch.cuda.device(g):
        bufs.append((torch.empty(total, dtype=torch.uint8, pin_memory=True), torch.empty(total, dtype=torch.uint8, device=f"cuda:{g}"),
                     torch.empty(total, dtype=torch.uint8, pin_memory=True), torch.empty(total, dtype=torch.uint8, device=f"cuda:{g}"),
                     torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)))


def run(gpus, h2d, d2h, reps=4):
    bar = threading.Barrier(len(gpus) + 1)
    def work(g):
        h, d, h2, d2, s1, s2 = bufs[g]
        torch.cuda.set_device(g)
        def once():
            for off in range(0, total, chunk):
                if h2d:
                    with torch.cuda.stream(s1):
                        d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h2[off:off + chunk].copy_(d2[off:off + chunk], non_blocking=True)
        once()
        torch.cuda.synchronize(g)
        bar.wait()
        for _ in range(reps):
            once()
        torch.cuda.synchronize(g)
        bar.wait()
    ts = [threading.Thread(target=work, args=(g,)) for g in gpus]
    [t.start() for t in ts]
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    [t.join() for t in ts]
    return total * reps / dt / 1e9  # per GPU, per direction


for k in sorted({1, 2, 4, n} & set(range(1, n + 1))):
    gpus = list(range(k))
    a, b, c = run(gpus, True, False), run(gpus, False, True), run(gpus, True, True)
    print(f"{k} GPU(s) at once, per GPU: H2D alone {a:5.1f}  D2H alone {b:5.1f}  both (each) {c:5.1f} GB/s   "
          f"aggregate: H2D {a * k:6.1f}  D2H {b * k:6.1f}  both {2 * c * k:6.1f} GB/s")
