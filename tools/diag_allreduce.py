"""Diagnostic (not shipped): where does the time of a 2-rank step go?"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t64 = torch.zeros(24, dtype=torch.float64, device=dev)
t32 = torch.zeros(24, dtype=torch.float32, device=dev)
big = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
def timeit(name, fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {e0.elapsed_time(e1)/n*1e3:.1f} us/iter device, {(time.perf_counter()-t0)/n*1e6:.1f} us wall", flush=True)
timeit("all_reduce f64[24]", lambda: dist.all_reduce(t64))
timeit("all_reduce f32[24]", lambda: dist.all_reduce(t32))
timeit("all_reduce f32[64Mi]", lambda: dist.all_reduce(big), 5)
def busy():
    big.mul_(1.0001)
    dist.all_reduce(t64)
timeit("kernel + all_reduce f64[24]", busy)
dist.destroy_process_group()
