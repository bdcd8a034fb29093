"""pinned-memory PCIe rates of the box: H2D alone, D2H alone, both at once (what bounds the end-to-end path)"""
import time
import torch

n = 256 << 20  # bytes per transfer
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n * reps / dt / 1e9


run(True, True, 2)
print("H2D alone   %.1f GB/s" % run(True, False))
print("D2H alone   %.1f GB/s" % run(False, True))
print("both at once %.1f GB/s each way" % run(True, True))
