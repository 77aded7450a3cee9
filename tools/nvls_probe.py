"""NVLS allreduce of the gradient bucket (gft_nvls_allreduce_sum) against NCCL: same sums, timing.
torchrun, 2+ GPUs of one NVSwitch domain."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gftorf_b200 import parallel
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = int(os.environ.get("P", "300000"))
g = torch.Generator().manual_seed(0)
params = {n: torch.zeros((P, k), device=dev) for n, k in parallel.PARAM_LAYOUT}
scal = [torch.zeros(1, device=dev), torch.zeros(1, device=dev)]
b = parallel.GradBucket(params, scal, symmetric=True)
def say(*a):
    if rank == 0: print(*a, flush=True)
say("bucket floats", b.flat.numel(), "NVLS path:", b._symm is not None)
x = torch.randn(b.flat.numel(), device=dev, generator=torch.Generator(dev).manual_seed(rank + 1))
ref = x.clone(); dist.all_reduce(ref)
b.flat.copy_(x)
b.allreduce()
torch.cuda.synchronize()
err = float((b.flat - ref).abs().max()); scale = float(ref.abs().max())
say("max |nvls - nccl| = %.3e (max |sum| %.3f)" % (err, scale))
assert err <= 1e-5 * scale
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
say("nvls  allreduce %.4f ms" % timeit(lambda: b.allreduce()))
say("two barriers alone %.4f ms" % timeit(lambda: (b._symm.barrier(channel=0), b._symm.barrier(channel=1))))
y = torch.randn(b.flat.numel(), device=dev)
say("nccl  allreduce %.4f ms" % timeit(lambda: dist.all_reduce(y)))
dist.destroy_process_group()
