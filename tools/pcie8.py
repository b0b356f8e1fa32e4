"""Pinned host<->device copy bandwidth of every rank at once (torchrun, or plain python for one GPU): H2D alone, D2H
alone, and both directions together on two streams -- the floor of the end-to-end arm, whose steps move the pictures
across PCIe once in each direction.  Prints one line per case: aggregate GB/s over the ranks and per rank."""
import os, time, torch, torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
for name, f, nbytes in (("H2D alone", h2d, n), ("D2H alone", d2h, n), ("H2D + D2H together (bytes of both)", both, 2 * n)):
    f(); barrier()
    t0 = time.perf_counter()
    for _ in range(6): f()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([6 * nbytes / dt / 1e9], device="cuda")
    if world > 1: dist.all_reduce(t)
    if int(os.environ.get("RANK", "0")) == 0:
        print("%-36s aggregate %6.1f GB/s over %d rank(s) (%.1f per rank)" % (name, t.item(), world, t.item() / world), flush=True)
if world > 1: dist.destroy_process_group()
