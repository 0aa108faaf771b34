"""PCIe checks for the e2e leg (GB/s, sum of both directions): pinned H2D alone, D2H alone, both at once - from one host thread on two
streams, and from two host threads the way the streamed e2e leg drives the library (each thread: big copy one way, small copy the other
way, small pageable copies and stream syncs in between)."""
import threading, time, torch
n = 1 << 28  # 256 MiB
small = 48 << 20
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_c = torch.empty(small, dtype=torch.uint8, pin_memory=True); h_d = torch.empty(small, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
d_c = torch.empty(small, dtype=torch.uint8, device="cuda"); d_d = torch.empty(small, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one_thread(h2d, d2h, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n * (int(h2d) + int(d2h)) / dt / 1e9
def two_threads(mixed, pageable, reps=6):
    def ta():
        with torch.cuda.stream(s1):
            for _ in range(reps):
                d_a.copy_(h_a, non_blocking=True)
                if pageable: _ = d_c[:8].cpu()
                if mixed: h_c.copy_(d_c, non_blocking=True)
                s1.synchronize()
    def tb():
        with torch.cuda.stream(s2):
            for _ in range(reps):
                if mixed: d_d.copy_(h_d, non_blocking=True)
                if pageable: _ = d_d[:8].cpu()
                h_b.copy_(d_b, non_blocking=True)
                s2.synchronize()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = threading.Thread(target=ta), threading.Thread(target=tb)
    a.start(); b.start(); a.join(); b.join()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * (2 * n + (2 * small if mixed else 0)) / dt / 1e9
one_thread(True, True, 1)
print("one thread: H2D alone %.1f, D2H alone %.1f, both %.1f GB/s" % (one_thread(True, False), one_thread(False, True), one_thread(True, True)))
print("two threads: big copies only %.1f, + small opposite copies %.1f, + small pageable reads %.1f GB/s" % (two_threads(False, False), two_threads(True, False), two_threads(True, True)))
