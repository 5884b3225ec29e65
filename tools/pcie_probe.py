"""PCIe probe (developer tool): H2D, D2H and both at once from pinned memory, to size the e2e floor."""
import torch, time
dev = torch.device("cuda:0")
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device=dev); d_b = torch.zeros(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both: %.1f ms for 1+1 GiB (sum alone %.1f ms) -> overlap %.0f%%" % (
    n / a / 1e9, n / b / 1e9, c * 1e3, (a + b) * 1e3, 100 * (a + b - c) / min(a, b)))
