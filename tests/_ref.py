"""TEST INFRASTRUCTURE: thin ctypes binding of oracle/_ref/libpointops_ref.so -- the UNMODIFIED
reference CUDA launchers compiled by oracle/build_ref.sh -- operating on torch CUDA tensors.
The reference launches on the legacy default stream and never checks errors, so every call is
bracketed by torch.cuda.synchronize()."""
from __future__ import annotations

import ctypes
from pathlib import Path

import torch

SO = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libpointops_ref.so"


def available() -> bool:
    return SO.exists() and torch.cuda.is_available()


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(SO))
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _sync():
    torch.cuda.synchronize()


def farthest_point_sampling(xyz, offset, new_offset):
    """Mirrors functions/sampling.py:8-23."""
    b = offset.shape[0]
    sizes = offset.clone()
    sizes[1:] -= offset[:-1]
    n_max = int(sizes.max().item())
    idx = torch.zeros(int(new_offset[-1].item()), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    _sync()
    lib().farthest_point_sampling_cuda_launcher(b, n_max, _p(xyz), _p(offset.int()), _p(new_offset.int()), _p(tmp), _p(idx))
    _sync()
    return idx


def knn_query(nsample, xyz, offset, new_xyz, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    o, no = offset.int(), new_offset.int()
    _sync()
    lib().knn_query_cuda_launcher(m, nsample, _p(xyz), _p(new_xyz), _p(o), _p(no), _p(idx), _p(dist2))
    _sync()
    return idx, dist2


def ball_query(nsample, max_radius, min_radius, xyz, offset, new_xyz, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    o, no = offset.int(), new_offset.int()
    _sync()
    lib().ball_query_cuda_launcher(m, nsample, ctypes.c_float(min_radius), ctypes.c_float(max_radius), _p(xyz),
                                   _p(new_xyz), _p(o), _p(no), _p(idx), _p(dist2))
    _sync()
    return idx, dist2


def random_ball_query(nsample, max_radius, min_radius, order, xyz, offset, new_xyz, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    o, no = offset.int(), new_offset.int()
    _sync()
    lib().random_ball_query_cuda_launcher(m, nsample, ctypes.c_float(min_radius), ctypes.c_float(max_radius),
                                          _p(order.int()), _p(xyz), _p(new_xyz), _p(o), _p(no), _p(idx), _p(dist2))
    _sync()
    return idx, dist2
