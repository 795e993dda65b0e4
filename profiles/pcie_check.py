import torch, time
x = torch.empty(50_000_000//8, dtype=torch.float64).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/10
print("H2D pinned GB/s", 0.05/dt)
t0=time.perf_counter()
for _ in range(10):
    x.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/10
print("D2H pinned GB/s", 0.05/dt)
y = torch.empty(50_000_000//8, dtype=torch.float64)
t0=time.perf_counter()
for _ in range(5):
    d.copy_(y)
torch.cuda.synchronize()
print("H2D pageable GB/s", 0.05/((time.perf_counter()-t0)/5))
