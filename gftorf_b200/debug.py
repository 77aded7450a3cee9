"""Debug accessor: typed views into the three opaque workspaces of gft_forward / gft_forward_views,
so the bit-exact tests can compare tiles_touched, sorted keys, point_list, tile ranges and n_contrib
with the reference's.  Offsets come from gft_workspace_layout_views
(include/gftorf.h); the layout is private to the library and may change between versions."""
import ctypes as C

import torch

from . import _capi


def decode_views(geom, binning, img, P, R, sizes):
    """`sizes`: [(W, H)] of the views of the batch, in call order.  Returns one dict per view; the
    binning arrays (`keys`, `point_list`) are the view's slice of the global lists and `ranges` is
    rebased to that slice, i.e. each dict reads like a single-view call."""
    V = len(sizes)
    lay = _capi.GftWorkspaceLayout()
    widths = (C.c_int * V)(*[int(w) for w, _ in sizes])
    heights = (C.c_int * V)(*[int(h) for _, h in sizes])
    _capi.lib().gft_workspace_layout_views(P, R, V, widths, heights, C.byref(lay))

    def v(buf, off, nbytes, dtype):
        return buf[off:off + nbytes].view(dtype)

    tiles = [((w + 15) // 16) * ((h + 15) // 16) for w, h in sizes]
    T_total, N_total = sum(tiles), sum(w * h for w, h in sizes)
    all_ranges = v(img, lay.img_ranges, 8 * T_total, torch.int32).view(T_total, 2)
    S = int(lay.img_sub_bins)
    all_counts = v(img, lay.img_tile_counts, 4 * T_total * S, torch.int32).view(T_total, S).sum(1, dtype=torch.int32)
    all_state = v(img, lay.img_state, 16 * N_total, torch.float32).view(N_total, 4)
    if R > 0:
        entries = v(binning, lay.bin_entries, 8 * R, torch.int64)
        point_list = v(binning, lay.bin_point_list, 4 * R, torch.int32)
    outs, t0, n0, r0 = [], 0, 0, 0
    for i, (W, H) in enumerate(sizes):
        out = {}
        T, N = tiles[i], W * H
        if P > 0:
            rec = v(geom, lay.geom_rec + i * P * 80, 80 * P, torch.float32).view(P, 20)
            clamped = v(geom, lay.geom_clamped + i * P * 4, 4 * P, torch.uint8).view(P, 4)
            tt = v(geom, lay.geom_tiles_touched + i * P * 4, 4 * P, torch.int32)
            out.update(
                rec=rec,
                means2D=rec[:, 0:2], extents=rec[:, 2:4], conic_opacity=rec[:, 4:8], rgb=rec[:, 8:11],
                dists=rec[:, 11], real_img_amp=rec[:, 12:19], ndc=rec[:, 19],
                depths=v(geom, lay.geom_depths + i * P * 4, 4 * P, torch.float32),
                tiles_touched=tt,
                # the reference's inclusive prefix sum (rasterizer_impl.cu:307); the library bins by
                # tile counts and never materialises it
                point_offsets=torch.cumsum(tt, 0, dtype=torch.int32),
                rect=v(geom, lay.geom_rect + i * P * 8, 8 * P, torch.int16).view(P, 4),
                cov3D=v(geom, lay.geom_cov3D, 24 * P, torch.float32).view(P, 6),
                clamped=clamped[:, 0:3], clamped_p=clamped[:, 3],
                pa=v(geom, lay.geom_pa + i * P * 8, 8 * P, torch.float32).view(P, 2),
            )
        state = all_state[n0:n0 + N]
        counts = all_counts[t0:t0 + T]
        Rv = int(counts.long().sum().item())
        rng = all_ranges[t0:t0 + T]
        # rebase the global ranges to this view's slice; empty tiles stay (0,0) as in the reference
        rng_local = torch.where((rng[:, 1] - rng[:, 0] > 0).unsqueeze(1), rng - r0, rng)
        out.update(
            final_T=state[:, 0], w_z_total=state[:, 1], w_z2_total=state[:, 2],
            n_contrib=state[:, 3].contiguous().view(torch.int32),
            ranges=rng_local, tile_counts=counts, num_rendered=Rv,
        )
        if R > 0 and Rv > 0:
            e = entries[r0:r0 + Rv]
            # the reference's key: (tile << 32) | float_bits(view_z); the tile of a sorted entry is
            # the range it lies in
            tile_of = torch.repeat_interleave(torch.arange(T, device=e.device), counts.long())
            out.update(
                point_list=point_list[r0:r0 + Rv],
                entries=e,
                keys=(tile_of << 32) | ((e >> 32) & 0xffffffff),
            )
        outs.append(out)
        t0 += T
        n0 += N
        r0 += Rv
    return outs


def decode_buffers(geom, binning, img, P, R, W, H):
    """Single-view call (rasterizer._C.rasterize_gaussians)."""
    return decode_views(geom, binning, img, P, R, [(W, H)])[0]
