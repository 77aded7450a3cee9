"""The one exchange step of data-parallel training at N ranks: the [P,91] fp32 gradient bucket summed
over the ranks — NCCL, the two-barrier NVLS kernel, the fused one-kernel NVLS allreduce (blocks /
unroll sweep), and torch's own symmetric-memory ops for reference.  torchrun, 2+ GPUs of one
NVSwitch domain:  torchrun --nproc-per-node N tools/exchange_probe.py [P]"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gftorf_b200 import parallel, train_ops

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 300000


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def timeit(fn, iters=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


params = {n: torch.zeros((P, k), device=dev) for n, k in parallel.PARAM_LAYOUT}
scal = [torch.zeros(1, device=dev), torch.zeros(1, device=dev)]
b = parallel.GradBucket(params, scal, symmetric=True)
say("bucket MB", b.nbytes() / 1e6, "symmetric:", b._symm is not None, "mode", b.mode)
res = {"P": P, "world": world, "MB": b.nbytes() / 1e6}
x = torch.randn(b.flat.numel(), device=dev, generator=torch.Generator(dev).manual_seed(rank + 1))
ref = x.clone(); dist.all_reduce(ref)
if b._symm is not None:
    for mode in ("nvls", "nvls_fused") + (("p2p_fused", "push_fused") if world <= 8 else ()):
        b.mode = mode
        b.flat.copy_(x)
        b.allreduce()
        torch.cuda.synchronize()
        err = float((b.flat - ref).abs().max()); scale = float(ref.abs().max())
        say(f"{mode}: max |x - nccl| = {err:.3e} (max |sum| {scale:.3f})")
        assert err <= 1e-5 * scale, mode
        res[mode] = timeit(lambda: b.allreduce())
        say(f"{mode:12s} {res[mode]:.4f} ms")
    hdl = b._symm
    lib = train_ops._lib()
    mc = int(hdl.multicast_ptr) + int(b.flat.data_ptr() - hdl.buffer_ptrs[hdl.rank])
    if world <= 8:
        b.mode = "p2p_fused"
        for shape in ((32, 2), (64, 2), (148, 1), (148, 2), (148, 4), (296, 2), (296, 4), (592, 2)):
            b.p2p_shape = shape
            b.flat.copy_(x)
            b.allreduce()
            torch.cuda.synchronize()
            assert float((b.flat - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), shape
            # all ranks must hold identical bits
            chk = b.flat.double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert float(lo) == float(hi), shape
            t = timeit(lambda: b.allreduce(), iters=20)
            res[f"p2p_b{shape[0]}_u{shape[1]}"] = t
            say(f"p2p blocks={shape[0]} unroll={shape[1]}: {t:.4f} ms")
    if world <= 8:
        b.mode = "push_fused"
        for nb in (32, 64, 148, 296, 592):
            b.push_blocks = nb
            b.flat.copy_(x)
            b.allreduce()
            torch.cuda.synchronize()
            assert float((b.flat - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), nb
            chk = b.flat.double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert float(lo) == float(hi), nb
            t = timeit(lambda: b.allreduce(), iters=20)
            res[f"push_b{nb}"] = t
            say(f"push blocks={nb}: {t:.4f} ms")
    for blocks in (16, 32, 48, 64, 96, 148, 296):
        for unroll in (2, 4, 8):
            def f():
                rc = lib.gft_nvls_allreduce_fused(C.c_void_p(mc), C.c_longlong(b.flat.numel()), hdl.rank, hdl.world_size,
                                                  C.c_void_p(int(hdl.signal_pad_ptrs_dev)), int(hdl.signal_pad_size) // 4,
                                                  blocks, unroll, C.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0
            t = timeit(f, iters=20)
            res[f"fused_b{blocks}_u{unroll}"] = t
            say(f"fused blocks={blocks} unroll={unroll}: {t:.4f} ms")
    say("two host barriers alone %.4f ms" % timeit(lambda: (hdl.barrier(channel=0), hdl.barrier(channel=1))))
    try:
        import torch.distributed._symmetric_memory as symm
        gname = dist.group.WORLD.group_name
        for op in ("multimem_all_reduce_", "two_shot_all_reduce_"):
            fn = getattr(torch.ops.symm_mem, op)
            res["torch_" + op] = timeit(lambda: fn(b.flat, "sum", gname))
            say("torch %-22s %.4f ms" % (op, res["torch_" + op]))
    except Exception as e:
        say("torch symm ops failed:", repr(e)[:200])
y = torch.randn(b.flat.numel(), device=dev)
res["nccl"] = timeit(lambda: dist.all_reduce(y))
say("nccl         %.4f ms" % res["nccl"])
b2 = parallel.GradBucket(params, scal, symmetric="auto")
say("autotune picked", b2.mode, b2.tuning)
res["autotune"] = {"mode": b2.mode, "ms": b2.tuning}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
