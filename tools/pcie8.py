"""Aggregate pinned H2D / D2H bandwidth with all ranks copying at once (torchrun)."""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, f in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): f()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([4 * n / dt / 1e9], device="cuda"); dist.all_reduce(t)
    if dist.get_rank() == 0: print("%s aggregate %.1f GB/s over %d ranks (%.1f per rank)" % (name, t.item(), dist.get_world_size(), t.item() / dist.get_world_size()), flush=True)
dist.destroy_process_group()
