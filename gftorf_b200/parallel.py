"""Multi-GPU host logic: view-sharded data parallelism for training and frame sharding for
render-only sweeps.

The reference is single-process, single-GPU and renders one view per iteration (train.py:159,
utils/general_utils.py:146); a rasterizer call is a pure function of (Gaussian parameters, one
camera), and loss gradients of different views simply add.  So the path shards by VIEW with the
Gaussian parameters replicated (364 B/Gaussian), and the only exchange step is one sum-allreduce of
the per-Gaussian gradients per iteration, plus a tiny sum/max exchange of the densification
statistics (scene/gaussian_model.py:648-654, train.py:443).  Render-only frames are independent:
no collective at all.

One process per GPU; torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

# parameter groups of scene/gaussian_model.py:247-272 as the rasterizer sees them, floats/Gaussian
PARAM_LAYOUT = (("means3D", 3), ("shs", 48), ("shs_p", 32), ("opacities", 1), ("scales", 3),
                ("rotations", 4))
FLOATS_PER_GAUSSIAN = sum(n for _, n in PARAM_LAYOUT)  # 91 -> 364 B


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Views of a multi-camera batch owned by `rank`: contiguous blocks, sizes differing by <= 1."""
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Frames of a render-only trajectory owned by `rank`: r, r+n, r+2n, ... (no collective)."""
    return list(range(rank, n_frames, world))


class ViewRunner:
    """Runs the rasterizer work of several views of ONE iteration concurrently: one host thread and
    one CUDA stream per view, forked from and joined to the caller's current stream.

    A view's forward + backward is a chain of latency-bound kernels (preprocess, five radix-sort
    passes, tile ranges: SMs mostly idle) around two large ones (blend forward / backward); two
    views in flight fill each other's gaps: -17 % per iteration at 640x480, -6 % at 1080p on B200
    (tools/overlap_probe.py).  The reference renders its colour and ToF views back to back on the
    default stream (gaussian_renderer/__init__.py:107-128); the views are independent until their
    gradients add, which `rasterizer._native_backward(..., accumulate="atomic")` does safely.

    Each callable runs under `torch.cuda.stream(its stream)`; ctypes releases the GIL while the
    library enqueues and waits, so the host threads overlap too."""

    def __init__(self, n_views: int, device=None):
        import threading
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_views)]
        self._jobs = [None] * n_views
        self._results = [None] * n_views
        self._go = [threading.Event() for _ in range(n_views)]
        self._done = [threading.Event() for _ in range(n_views)]
        self._stop = False
        self._threads = [threading.Thread(target=self._loop, args=(i,), daemon=True) for i in range(n_views)]
        for t in self._threads:
            t.start()

    def _loop(self, i):
        torch.cuda.set_device(self.device)
        while True:
            self._go[i].wait()
            self._go[i].clear()
            if self._stop:
                return
            try:
                with torch.cuda.stream(self.streams[i]):
                    self._results[i] = (True, self._jobs[i]())
            except BaseException as e:      # handed to the caller's thread
                self._results[i] = (False, e)
            self._done[i].set()

    def run(self, fns):
        """fns: one callable per view.  Returns their results in order; re-raises the first error."""
        if len(fns) > len(self.streams):
            raise RuntimeError("ViewRunner was built for fewer views")
        main = torch.cuda.current_stream(self.device)
        for i, fn in enumerate(fns):
            self.streams[i].wait_stream(main)
            self._jobs[i] = fn
            self._go[i].set()
        out = []
        for i in range(len(fns)):
            self._done[i].wait()
            self._done[i].clear()
            main.wait_stream(self.streams[i])
        for i in range(len(fns)):
            ok, val = self._results[i]
            self._jobs[i] = self._results[i] = None
            if not ok:
                raise val
            out.append(val)
        return out

    def close(self):
        self._stop = True
        for e in self._go:
            e.set()


class GradBucket:
    """One flat fp32 buffer holding the gradient of every Gaussian parameter (+ the two scalar ToF
    offsets).  `attach()` makes each leaf's .grad a VIEW into the buffer, so autograd accumulates
    the gradients of all local views in place and the collective runs on the buffer as it stands —
    no pack or copy step between the last backward and the allreduce."""

    def __init__(self, params: Dict[str, torch.Tensor], scalars: Sequence[torch.Tensor] = (),
                 symmetric=False, group=None):
        """`symmetric`: False — an ordinary tensor, exchanged with NCCL; True — symmetric memory
        bound to an NVLink multicast object, exchanged by the library's one-kernel NVLS allreduce
        (`gft_nvls_allreduce_fused`); "auto" — symmetric memory, and `autotune()` (called here)
        times NCCL against the NVLS kernels on this very buffer and keeps the fastest."""
        self.names = [n for n, _ in PARAM_LAYOUT if n in params]
        self.params = params
        self.scalars = list(scalars)
        ref = params[self.names[0]]
        sizes = [params[n].numel() for n in self.names] + [s.numel() for s in self.scalars]
        self.offsets, cur = [], 0
        for n in sizes:
            self.offsets.append(cur)
            cur += (n + 3) // 4 * 4  # 16-byte aligned slices
        self.sizes = sizes
        self._symm = None
        self.fused_shape = (64, 2)   # (blocks, unroll) of the one-kernel NVLS allreduce: best of the sweep at 4 and 8 GPUs
        self.p2p_shape = (148, 2)    # (blocks, unroll) of the one-kernel peer-to-peer allreduce
        self.push_blocks = 148       # blocks of the stores-only peer-to-peer allreduce
        self._scratch = None         # its symmetric scratch buffer (allocated on first use)
        self._group = group
        self.mode = "nccl"
        self.tuning = None
        if symmetric:
            self.flat = self._alloc_symmetric(cur, ref.device, group)
            if self._symm is not None:
                self.mode = "nvls_fused"
                if symmetric == "auto":
                    self.autotune(group)
        else:
            self.flat = torch.zeros(cur, dtype=torch.float32, device=ref.device)

    def _alloc_symmetric(self, n, device, group):
        """The bucket in symmetric memory bound to an NVLink multicast object, so the exchange can
        run through the switch.  Collective: every rank of `group` must construct its bucket at
        the same point.  Falls back to an ordinary tensor (and NCCL) when the platform offers no
        multicast."""
        try:
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else dist.group.WORLD
            try:        # room for one signal word per (block, peer) of the fused kernel's barriers
                if symm.get_signal_pad_size() < 65536:
                    symm.set_signal_pad_size(65536)
            except Exception:
                pass
            flat = symm.empty(n, dtype=torch.float32, device=device)
            hdl = symm.rendezvous(flat, g.group_name)
            flat.zero_()
            if int(hdl.multicast_ptr) != 0:
                self._symm = hdl
            return flat
        except Exception:
            self._symm = None
            return torch.zeros(n, dtype=torch.float32, device=device)

    def _scratch_handle(self):
        """Second symmetric buffer for the stores-only exchange: world x ceil(n / 4 / world) float4
        per rank.  Collective on first use (every rank reaches it at the same call)."""
        if self._scratch is None:
            import torch.distributed._symmetric_memory as symm
            g = self._group if self._group is not None else dist.group.WORLD
            w = self._symm.world_size
            n4 = self.flat.numel() // 4
            per = (n4 + w - 1) // w
            t = symm.empty(per * w * 4, dtype=torch.float32, device=self.flat.device)
            self._scratch = (t, symm.rendezvous(t, g.group_name))
        return self._scratch[1]

    def autotune(self, group=None, iters=8):
        """Times the available exchanges on the bucket (contents preserved) and keeps the fastest
        (max over ranks, so every rank picks the same).  Collective."""
        if self._symm is None or not dist.is_initialized():
            return
        keep = self.flat.clone()
        res = {}
        shapes = {"nvls_fused": (64, 2), "nvls_fused_148": (148, 2), "nvls_fused_u4": (64, 4)}
        p2p_shapes = {"p2p_fused": (148, 2), "p2p_fused_64": (64, 2), "p2p_fused_296u4": (296, 4)}
        keys = ["nccl", "nvls", "nvls_fused", "nvls_fused_148", "nvls_fused_u4"]
        push_shapes = {"push_fused": 148, "push_fused_64": 64, "push_fused_296": 296}
        if self._symm.world_size <= 4:        # the peer-to-peer exchange moves more bytes than the switch from N = 4 on
            keys += list(p2p_shapes)          # (mode "push_fused", stores only, measured 0.215 vs 0.207 ms at N = 2: not tried here)
        for key in keys:
            mode = "nvls_fused" if key in shapes else ("p2p_fused" if key in p2p_shapes else
                                                       ("push_fused" if key in push_shapes else key))
            self.mode = mode
            if key in push_shapes:
                self.push_blocks = push_shapes[key]
            if key in shapes:
                self.fused_shape = shapes[key]
            if key in p2p_shapes:
                self.p2p_shape = p2p_shapes[key]
            try:
                for _ in range(3):
                    self.allreduce(group)
                torch.cuda.synchronize()
                dist.barrier(group)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    self.allreduce(group)
                b.record()
                torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / iters], device=self.flat.device)
            except Exception:
                t = torch.tensor([float("inf")], device=self.flat.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            res[key] = float(t)
        best = min(res, key=res.get)
        self.mode = "nvls_fused" if best in shapes else ("p2p_fused" if best in p2p_shapes else
                                                        ("push_fused" if best in push_shapes else best))
        self.push_blocks = push_shapes.get(best, 148)
        self.fused_shape = shapes.get(best, (64, 2))
        self.p2p_shape = p2p_shapes.get(best, (148, 2))
        self.tuning = {k: round(v, 4) for k, v in res.items()}
        self.flat.copy_(keep)

    def views(self):
        tensors = [self.params[n] for n in self.names] + self.scalars
        return [self.flat[o:o + n].view_as(t) for o, n, t in zip(self.offsets, self.sizes, tensors)]

    def attach(self):
        tensors = [self.params[n] for n in self.names] + self.scalars
        for t, v in zip(tensors, self.views()):
            t.grad = v
        return self

    def grad_out(self):
        """The bucket's slices keyed the way `rasterizer._native_backward(..., grad_out=...)` wants
        them: the backward kernel then ADDS each view's gradients straight into the bucket
        (no autograd accumulation pass, no copy)."""
        out = dict(zip(self.names, self.views()[:len(self.names)]))
        sc = self.views()[len(self.names):]
        if len(sc) >= 2:
            out["phase_offset"], out["dc_offset"] = sc[0], sc[1]
        return out

    def zero(self):
        self.flat.zero_()

    def nbytes(self):
        return self.flat.numel() * 4

    def allreduce(self, group=None, average: bool = False, async_op: bool = False):
        """Sum (or mean) over ranks.  No-op in a single process."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if self._symm is not None and not async_op and self.mode != "nccl":
            import ctypes as C
            from . import train_ops
            hdl = self._symm
            lib = train_ops._lib()
            mc = int(hdl.multicast_ptr) + int(self.flat.data_ptr() - hdl.buffer_ptrs[hdl.rank])
            stream = C.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
            if self.mode == "push_fused":
                # ONE kernel, posted stores only: push foreign slices into the owners' scratch, barrier,
                # owners sum and push the sums into every bucket
                sh = self._scratch_handle()
                ptrs = (C.c_void_p * hdl.world_size)(*[int(p) for p in hdl.buffer_ptrs])
                sptrs = (C.c_void_p * hdl.world_size)(*[int(p) for p in sh.buffer_ptrs])
                with torch.cuda.device(self.flat.device):
                    rc = lib.gft_push_allreduce_fused(ptrs, C.c_longlong(int(self.flat.data_ptr() - hdl.buffer_ptrs[hdl.rank])),
                                                      sptrs, C.c_longlong(self.flat.numel()), hdl.rank, hdl.world_size,
                                                      C.c_void_p(int(hdl.signal_pad_ptrs_dev)),
                                                      int(hdl.signal_pad_size) // 4, int(self.push_blocks), stream)
                train_ops._check(rc, "gft_push_allreduce_fused")
            elif self.mode == "p2p_fused":
                # ONE kernel, plain peer-to-peer accesses: this rank sums its slice from every rank's
                # buffer and stores the sum into every buffer (fewer bytes than the multicast path at N = 2)
                ptrs = (C.c_void_p * hdl.world_size)(*[int(p) for p in hdl.buffer_ptrs])
                with torch.cuda.device(self.flat.device):
                    rc = lib.gft_p2p_allreduce_fused(ptrs, C.c_longlong(int(self.flat.data_ptr() - hdl.buffer_ptrs[hdl.rank])),
                                                     C.c_longlong(self.flat.numel()), hdl.rank, hdl.world_size,
                                                     C.c_void_p(int(hdl.signal_pad_ptrs_dev)),
                                                     int(hdl.signal_pad_size) // 4, int(self.p2p_shape[0]),
                                                     int(self.p2p_shape[1]), stream)
                train_ops._check(rc, "gft_p2p_allreduce_fused")
            elif self.mode == "nvls_fused":
                # ONE kernel: barrier over the signal pads, multimem.ld_reduce + multimem.st on this
                # rank's slice, barrier
                with torch.cuda.device(self.flat.device):
                    rc = lib.gft_nvls_allreduce_fused(C.c_void_p(mc), C.c_longlong(self.flat.numel()), hdl.rank,
                                                      hdl.world_size, C.c_void_p(int(hdl.signal_pad_ptrs_dev)),
                                                      int(hdl.signal_pad_size) // 4, int(self.fused_shape[0]),
                                                      int(self.fused_shape[1]), stream)
                train_ops._check(rc, "gft_nvls_allreduce_fused")
            else:
                # host-launched barrier, one kernel, barrier — all on the current stream
                hdl.barrier(channel=0)
                with torch.cuda.device(self.flat.device):
                    rc = lib.gft_nvls_allreduce_sum(C.c_void_p(mc), C.c_longlong(self.flat.numel()), hdl.rank,
                                                    hdl.world_size, stream)
                train_ops._check(rc, "gft_nvls_allreduce_sum")
                hdl.barrier(channel=1)
            if average:
                self.flat.div_(hdl.world_size)
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            if async_op:
                work.wait()
                work = None
            self.flat.div_(dist.get_world_size(group))
        return work


class DensifyStats:
    """Per-Gaussian densification statistics (scene/gaussian_model.py:648-654, train.py:441-445).
    The gradient NORM is taken per view, so each rank accumulates its own views and the ranks
    exchange sums (accum, denom) and a max (max_radii2D)."""

    def __init__(self, P: int, device):
        self.sums = torch.zeros((P, 2), dtype=torch.float32, device=device)   # grad-norm*pixels, pixels
        self.max_radii2D = torch.zeros((P,), dtype=torch.float32, device=device)

    def update(self, viewspace_grad: torch.Tensor, radii: torch.Tensor, pixels: torch.Tensor):
        vis = radii > 0
        self.max_radii2D[vis] = torch.max(self.max_radii2D[vis], radii[vis].float())
        pix = pixels.view(-1)
        norm = torch.norm(viewspace_grad[:, :2], dim=-1)
        self.sums[:, 0] += torch.where(vis, norm * pix, torch.zeros_like(norm))
        self.sums[:, 1] += torch.where(vis, pix, torch.zeros_like(pix))

    def allreduce(self, group=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=group)


def train_views(rasterize_view, views: Sequence, bucket: GradBucket, group=None,
                stats: Optional[DensifyStats] = None):
    """One data-parallel iteration over this rank's views.

    `rasterize_view(view)` runs forward + loss + backward for one camera through the public
    autograd surface and returns (viewspace_points, radii, pixels) for the densification
    statistics; gradients land in the bucket through the attached .grad views.  Afterwards the
    bucket (and the statistics) are reduced over the ranks."""
    bucket.zero()
    for v in views:
        out = rasterize_view(v)
        if stats is not None and out is not None:
            vsp, radii, pixels = out
            stats.update(vsp.grad, radii, pixels)
    bucket.allreduce(group=group)
    if stats is not None:
        stats.allreduce(group=group)
