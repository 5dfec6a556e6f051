import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda"); h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
print("D2H 1 GiB alone: %.2f ms" % t(lambda: h.copy_(d, non_blocking=True)))
print("H2D 1 GiB alone: %.2f ms" % t(lambda: d.copy_(h, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
print("D2H 1 GiB + concurrent H2D 0.5 GiB: %.2f ms" % t(both))
