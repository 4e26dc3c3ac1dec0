"""All-gather / all-reduce bandwidth probe (torchrun): what the NCCL plumbing delivers on this box."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (64, 512, 2048):
    n = mb * 1024 * 1024 // 8
    buf = torch.zeros(world * n, dtype=torch.int64, device="cuda")
    x = torch.zeros(n, dtype=torch.int64, device="cuda")
    for name, fn in (("all_gather_into_tensor(in place)", lambda: dist.all_gather_into_tensor(buf, buf[rank * n:(rank + 1) * n])),
                     ("all_gather_into_tensor", lambda: dist.all_gather_into_tensor(buf, x)),
                     ("all_reduce", lambda: dist.all_reduce(x))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): fn()
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        if rank == 0:
            recv = (world - 1) * mb / 1024 if "gather" in name else 2 * (world - 1) / world * mb / 1024
            print(f"{name:34s} {mb:5d} MB/rank  {ms:8.2f} ms   {recv / (ms / 1000):7.1f} GB/s per-rank wire traffic", flush=True)
dist.destroy_process_group()
