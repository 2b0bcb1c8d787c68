"""torchrun --nproc-per-node N tools/pcie_probe_multi.py : host<->device copy bandwidth of every rank with all N ranks
copying AT THE SAME TIME (pinned memory, 1 GiB H2D / 512 MiB D2H per repetition), to tell a per-link limit from a shared
host-side one (memory bandwidth, PCIe switch uplinks).  Rank 0 prints one line per mode."""
import os, time
import torch, torch.distributed as td

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = td.get_rank(), td.get_world_size()
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


def bw(fn, nbytes, reps=4):
    fn(); torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


for name, fn, nb in (("H2D", h2d, n), ("D2H", d2h, n // 2), ("H2D+D2H (H2D bytes/time)", both, n)):
    v = torch.tensor([bw(fn, nb)], device="cuda")
    allv = [torch.zeros_like(v) for _ in range(world)]
    td.all_gather(allv, v)
    if rank == 0:
        vals = [float(a.item()) for a in allv]
        print(f"{world} ranks concurrently, {name}: per GPU " + " ".join(f"{x:5.1f}" for x in vals) + f" GB/s | sum {sum(vals):.1f} GB/s", flush=True)
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    aff = pynvml.nvmlDeviceGetCpuAffinity(h, 4)
    info = f"rank {rank}: pci {pynvml.nvmlDeviceGetPciInfo(h).busId} cpu-affinity words {[hex(a) for a in aff]} host cpus {os.cpu_count()}"
except Exception as ex:
    info = f"rank {rank}: nvml unavailable ({ex!r})"
box = [None] * world
td.all_gather_object(box, info)
if rank == 0:
    print("\n".join(box), flush=True)
td.destroy_process_group()
