import torch, time
x = torch.empty(1 << 27, dtype=torch.int64).pin_memory()   # 1 GiB
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("H2D pinned 1GiB: %.1f GB/s" % (x.numel() * 8 / dt / 1e9))
h = torch.empty(1 << 24, dtype=torch.int64).pin_memory()
t0 = time.perf_counter()
for _ in range(5): h.copy_(d[: 1 << 24], non_blocking=True)
torch.cuda.synchronize()
print("D2H pinned 128MiB: %.1f GB/s" % (h.numel() * 8 * 5 / (time.perf_counter() - t0) / 1e9))
