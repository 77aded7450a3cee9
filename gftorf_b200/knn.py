"""distCUDA2 — mean squared distance to the 3 nearest other points.

Mirrors `simple_knn._C.distCUDA2` of the reference (submodules/simple-knn/ext.cpp:15-17,
spatial.cu:15-26): takes a float CUDA tensor [P,3], returns a float tensor [P].  Used once at
initialisation by scene/gaussian_model.py:194-198.
"""
import ctypes as C

import torch

from . import _capi


def distCUDA2(points):
    if not points.is_cuda:
        raise RuntimeError("gftorf_b200 has no CPU path: distCUDA2 needs a CUDA tensor")
    lib = _capi.lib()
    P = int(points.shape[0])
    pts = points.contiguous().float()
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)  # spatial.cu:20
    if P == 0:
        return means
    ws = torch.empty(lib.gft_dist2_workspace_bytes(P), dtype=torch.uint8, device=points.device)
    stream = torch.cuda.current_stream(points.device).cuda_stream
    with torch.cuda.device(points.device):
        rc = lib.gft_dist2(pts.data_ptr(), P, means.data_ptr(), ws.data_ptr(), C.c_void_p(stream))
    if rc < 0:
        raise RuntimeError("gft_dist2 failed: " + lib.gft_last_error().decode())
    return means
