"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: view/frame sharding, the flat gradient
bucket whose views ARE the leaves' .grad, its sum-allreduce, and the densification statistics
exchange.  The data-path kernels are not involved; gradients here come from a small
differentiable stand-in so autograd's in-place accumulation into the bucket is exercised."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gftorf_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_params(P, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = dict(means3D=(P, 3), shs=(P, 16, 3), shs_p=(P, 16, 2), opacities=(P, 1), scales=(P, 3),
                  rotations=(P, 4))
    return {k: torch.rand(s, generator=g).requires_grad_(True) for k, s in shapes.items()}


def _view_loss(params, scalars, view):
    # any differentiable function of all parameters that differs per view
    w = float(view + 1)
    return sum(((t * w) ** 2).sum() for t in params.values()) + w * (scalars[0] * 3 + scalars[1] * 5).sum()


def _worker(rank, world, port, n_views, P, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _make_params(P)
        scalars = [torch.zeros(1, requires_grad=True), torch.zeros(1, requires_grad=True)]
        bucket = parallel.GradBucket(params, scalars).attach()
        stats = parallel.DensifyStats(P, "cpu")
        mine = parallel.shard_views(n_views, rank, world)

        def rasterize_view(v):
            _view_loss(params, scalars, v).backward()
            vsp = torch.zeros((P, 3))
            vsp.grad = torch.full((P, 3), float(v + 1))
            radii = torch.full((P,), v + 1, dtype=torch.int32)
            radii[::2] = 0
            return vsp, radii, torch.full((P, 1), 2.0)

        parallel.train_views(rasterize_view, mine, bucket, stats=stats)
        # .grad views alias the bucket: the reduced values are visible through the leaves
        assert params["means3D"].grad.data_ptr() == bucket.flat.data_ptr()
        out[rank] = dict(flat=bucket.flat.clone(), sums=stats.sums.clone(), maxr=stats.max_radii2D.clone(),
                         views=mine, means_grad=params["means3D"].grad.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [2, 5, 8])
def test_view_sharded_gradients_equal_the_unsharded_sum(n_views):
    P, world = 37, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_views, P, out), nprocs=world, join=True)
    # single-process reference over the whole batch
    params = _make_params(P)
    scalars = [torch.zeros(1, requires_grad=True), torch.zeros(1, requires_grad=True)]
    bucket = parallel.GradBucket(params, scalars).attach()
    for v in range(n_views):
        _view_loss(params, scalars, v).backward()
    assert sorted(out[0]["views"] + out[1]["views"]) == list(range(n_views))
    for r in range(world):
        assert torch.allclose(out[r]["flat"], bucket.flat, rtol=1e-6, atol=1e-6)
        assert torch.allclose(out[r]["means_grad"], params["means3D"].grad, rtol=1e-6)
    assert torch.equal(out[0]["flat"], out[1]["flat"])          # identical on every rank
    # densification statistics: sums add over views, radii take the max
    vis = torch.arange(P) % 2 == 1
    exp_pix = torch.where(vis, torch.tensor(2.0 * n_views), torch.tensor(0.0))
    assert torch.allclose(out[0]["sums"][:, 1], exp_pix)
    assert torch.equal(out[0]["maxr"][vis], torch.full((int(vis.sum()),), float(n_views)))
    assert float(out[0]["maxr"][~vis].abs().sum()) == 0.0
    exp_norm = sum(2.0 * (2 * (v + 1) ** 2) ** 0.5 for v in range(n_views))
    assert torch.allclose(out[0]["sums"][vis, 0], torch.full((int(vis.sum()),), exp_norm), rtol=1e-5)


def test_sharding_partitions():
    for n, w in ((8, 1), (8, 2), (8, 4), (8, 8), (5, 2), (3, 4), (120, 8)):
        views = [parallel.shard_views(n, r, w) for r in range(w)]
        assert sorted(sum(views, [])) == list(range(n))
        assert max(len(v) for v in views) - min(len(v) for v in views) <= 1
        frames = [parallel.shard_frames(n, r, w) for r in range(w)]
        assert sorted(sum(frames, [])) == list(range(n))
        for r in range(w):
            assert all(f % w == r for f in frames[r])


def test_bucket_layout():
    params = _make_params(11)
    b = parallel.GradBucket(params)
    assert parallel.FLOATS_PER_GAUSSIAN == 91
    assert b.flat.numel() >= 11 * 91
    for v in b.views():
        assert v.data_ptr() % 16 == 0
    b.attach()
    params["shs"].sum().backward()
    assert float(b.flat.sum()) == params["shs"].numel()
    assert b.allreduce() is None        # single process: no-op
