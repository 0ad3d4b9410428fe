"""Development probe: host-to-device bandwidth from pinned memory, one stream vs two (the e2e ceiling of bench.py)."""
import torch, time
n = 65536*4096
xh = torch.empty(n, dtype=torch.float32).pin_memory(); th = torch.empty(n, dtype=torch.float32).pin_memory()
x = torch.empty(n, dtype=torch.float32, device="cuda"); t = torch.empty(n, dtype=torch.float32, device="cuda")
s2 = torch.cuda.Stream()
def one():
    x.copy_(xh, non_blocking=True); t.copy_(th, non_blocking=True)
def two():
    with torch.cuda.stream(s2):
        t.copy_(th, non_blocking=True)
    x.copy_(xh, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
for name, fn in (("one stream", one), ("two streams", two), ("one stream", one), ("two streams", two)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {2*n*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms for 2 GiB)")
