"""All-reduce bandwidth probe (experiment): torchrun --nproc-per-node N scripts/nccl_probe.py"""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dist.get_world_size()
for gb in (0.25, 1.0, 10.0):
    x = torch.ones(int(gb * 1e9 / 4), dtype=torch.float32, device="cuda")
    for _ in range(2): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): dist.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    if dist.get_rank() == 0:
        print("all_reduce %.2f GB fp32 x%d ranks: %.2f ms  algbw %.0f GB/s  busbw %.0f GB/s" % (gb, w, ms, gb / ms * 1e3, gb / ms * 1e3 * 2 * (w - 1) / w), flush=True)
    del x
dist.destroy_process_group()
