import torch, time
n = 512*512*513
h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
print("H2D GB/s", gb / t(lambda: d.copy_(h, non_blocking=True)))
print("D2H GB/s", gb / t(lambda: h.copy_(d, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("duplex: each direction GB/s", gb / t(both))
