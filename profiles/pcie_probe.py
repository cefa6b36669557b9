"""Raw pinned-memory PCIe bandwidth of the box (what bounds bench.py's e2e arm)."""
import time
import torch
n = 369 * 1000 * 1000 // 4
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
h_in = torch.empty(n // 2, dtype=torch.float32).pin_memory()
d_out = torch.empty(n, dtype=torch.float32, device="cuda")
d_in = torch.empty(n // 2, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if d2h:
            with torch.cuda.stream(s1): h_out.copy_(d_out, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2): d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
for name, a, b in (("d2h only", False, True), ("h2d only", True, False), ("both", True, True)):
    run(a, b, 2); t = run(a, b)
    print("%s: %.2f ms  d2h %.1f GB/s  h2d %.1f GB/s" % (name, 1e3 * t, (n * 4 / t / 1e9) if b else 0, (n * 2 / t / 1e9) if a else 0))
