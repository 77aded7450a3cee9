"""How fast can the [P,91] gradient bucket be summed over the ranks?  NCCL on a plain tensor, NCCL on
a registered (ncclMemAlloc) buffer, and the symmetric-memory NVLS paths.  torchrun, 2+ GPUs."""
import os, time, sys
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 300000 * 91 + 8
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
def say(*a):
    if rank == 0: print(*a, flush=True)
x = torch.randn(n, device=dev)
say("nccl plain            %.4f ms" % timeit(lambda: dist.all_reduce(x)))
try:
    backend = dist.group.WORLD._get_backend(dev)
    pool = torch.cuda.MemPool(backend.mem_allocator)
    with torch.cuda.use_mem_pool(pool):
        y = torch.randn(n, device=dev)
    backend.register_mem_pool(pool)
    say("nccl registered pool  %.4f ms" % timeit(lambda: dist.all_reduce(y)))
except Exception as e:
    say("registered pool failed:", repr(e)[:300])
try:
    import torch.distributed._symmetric_memory as symm
    gname = dist.group.WORLD.group_name
    z = symm.empty(n, dtype=torch.float32, device=dev)
    h = symm.rendezvous(z, gname)
    z.normal_()
    say("symm handle: multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None, "world", h.world_size)
    for op in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
        try:
            f = getattr(torch.ops.symm_mem, op)
            say("symm %-22s %.4f ms" % (op, timeit(lambda: f(z, "sum", gname))))
        except Exception as e:
            say("symm", op, "failed:", repr(e)[:200])
    # correctness of multimem path vs nccl
    z.copy_(torch.arange(n, device=dev, dtype=torch.float32) % 7 + rank)
    ref = z.clone(); dist.all_reduce(ref)
    torch.ops.symm_mem.multimem_all_reduce_(z, "sum", gname)
    say("multimem == nccl:", bool(torch.equal(z, ref)))
except Exception as e:
    say("symmetric memory failed:", repr(e)[:400])
dist.destroy_process_group()
