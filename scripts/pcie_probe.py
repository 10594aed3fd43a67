import torch, time
dev = torch.device("cuda:0")
n = 512 * 1024 * 1024  # 2 GiB fp32
h = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev); d2 = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, k=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
gb = n * 4 / 1e9
print("H2D %.1f GB/s" % (gb / t(lambda: d.copy_(h, non_blocking=True))))
print("D2H %.1f GB/s" % (gb / t(lambda: h2.copy_(d2, non_blocking=True))))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("H2D+D2H concurrently: %.1f GB/s each" % (gb / t(both)))
