"""PCIe check for the e2e leg: pinned H2D alone, D2H alone, and both at once on two streams (GB/s)."""
import torch, time
n = 1 << 29  # 512 MiB
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n * (int(h2d) + int(d2h)) / dt / 1e9
run(True, True, 1)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both %.1f GB/s (sum of the two directions)" % (run(True, False), run(False, True), run(True, True)))
