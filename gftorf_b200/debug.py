"""Debug accessor: typed views into the three opaque workspaces of gft_forward, so the bit-exact
tests can compare tiles_touched, sorted keys, point_list, tile ranges and n_contrib with the
reference's.  Offsets come from gft_workspace_layout (include/gftorf.h);
the layout is private to the library and may change between versions."""
import ctypes as C

import torch

from . import _capi


def decode_buffers(geom, binning, img, P, R, W, H):
    lay = _capi.GftWorkspaceLayout()
    _capi.lib().gft_workspace_layout(P, R, W, H, C.byref(lay))
    N = W * H
    T = ((W + 15) // 16) * ((H + 15) // 16)

    def v(buf, off, nbytes, dtype):
        return buf[off:off + nbytes].view(dtype)

    out = {}
    if P > 0:
        rec = v(geom, lay.geom_rec, 80 * P, torch.float32).view(P, 20)
        clamped = v(geom, lay.geom_clamped, 4 * P, torch.uint8).view(P, 4)
        out.update(
            rec=rec,
            means2D=rec[:, 0:2], extents=rec[:, 2:4], conic_opacity=rec[:, 4:8], rgb=rec[:, 8:11],
            dists=rec[:, 11], real_img_amp=rec[:, 12:19], ndc=rec[:, 19],
            depths=v(geom, lay.geom_depths, 4 * P, torch.float32),
            tiles_touched=v(geom, lay.geom_tiles_touched, 4 * P, torch.int32),
            point_offsets=v(geom, lay.geom_point_offsets, 4 * P, torch.int32),
            rect=v(geom, lay.geom_rect, 8 * P, torch.int16).view(P, 4),
            cov3D=v(geom, lay.geom_cov3D, 24 * P, torch.float32).view(P, 6),
            clamped=clamped[:, 0:3], clamped_p=clamped[:, 3],
            pa=v(geom, lay.geom_pa, 8 * P, torch.float32).view(P, 2),
        )
    state = v(img, lay.img_state, 16 * N, torch.float32).view(N, 4)
    out.update(
        final_T=state[:, 0], w_z_total=state[:, 1], w_z2_total=state[:, 2],
        n_contrib=state[:, 3].contiguous().view(torch.int32),
        ranges=v(img, lay.img_ranges, 8 * T, torch.int32).view(T, 2),
    )
    if R > 0:
        # The library's sort key is (tile << depth_bits) | (float_bits(view_z) - depth_base) with
        # the format recorded in words 2, 3 of the geometry header; `keys` is rebuilt in the
        # reference's format (tile << 32) | float_bits(view_z) for the bit-exact comparisons,
        # `keys_compact` is what the sort actually saw.
        hdr = v(geom, 0, 16, torch.int32)
        depth_bits, depth_base = int(hdr[2]), int(hdr[3]) & 0xffffffff
        ck = v(binning, lay.bin_keys, 8 * R, torch.int64)
        if depth_bits >= 32:
            keys = ck
        else:
            keys = ((ck >> depth_bits) << 32) | ((ck & ((1 << depth_bits) - 1)) + depth_base)
        out.update(
            point_list=v(binning, lay.bin_point_list, 4 * R, torch.int32),
            keys=keys, keys_compact=ck, key_depth_bits=depth_bits,
        )
    return out
