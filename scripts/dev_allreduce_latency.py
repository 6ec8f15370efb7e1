"""What SURVEY 8e grain 2 (column blocks of ONE matrix on several GPUs) would pay per greedy step: back-to-back NCCL
all-reduces of a partial pair-counter slab.  torchrun --nproc-per-node N scripts/dev_allreduce_latency.py"""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl')
for nbytes in (4 << 10, 256 << 10, 4 << 20, 7 << 20):
    x = torch.ones(nbytes // 4, dtype=torch.int32, device='cuda')
    for _ in range(50):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    n = 2000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(n):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'all_reduce int32 sum, {world} ranks, {nbytes >> 10} KB: {1e3 * t.item() / n:.1f} us each ({n} back to back)', flush=True)
dist.destroy_process_group()
