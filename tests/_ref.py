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


# ---- the gather / scatter launchers (grouping, interpolation, subtraction, aggregation, scatter-attention) --------------
# Output / gradient buffers are zero-initialised exactly as the reference's Python wrappers allocate them
# (functions/grouping.py:14,26, interpolation.py:47,57, subtraction.py, aggregation.py, attention.py use
# torch.cuda.FloatTensor(...).zero_()): the backward kernels accumulate with atomicAdd.
def _z(shape, dev):
    return torch.zeros(shape, dtype=torch.float32, device=dev)


def grouping_forward(inp, idx):
    """grouping_forward_cuda_launcher (src/grouping/grouping_cuda_kernel.cu): out[m, s, :] = input[idx[m, s], :]."""
    (m, ns), c = idx.shape, inp.shape[1]
    out = _z((m, ns, c), inp.device)
    _sync()
    lib().grouping_forward_cuda_launcher(m, ns, c, _p(inp), _p(idx), _p(out))
    _sync()
    return out


def grouping_backward(grad_out, idx, n):
    m, ns, c = grad_out.shape
    gi = _z((n, c), grad_out.device)
    _sync()
    lib().grouping_backward_cuda_launcher(m, ns, c, _p(grad_out), _p(idx), _p(gi))
    _sync()
    return gi


def interpolation_forward(inp, idx, weight):
    """interpolation_forward_cuda_launcher: out[n, :] = sum_k weight[n, k] * input[idx[n, k], :]."""
    (n, k), c = idx.shape, inp.shape[1]
    out = _z((n, c), inp.device)
    _sync()
    lib().interpolation_forward_cuda_launcher(n, c, k, _p(inp), _p(idx), _p(weight), _p(out))
    _sync()
    return out


def interpolation_backward(grad_out, idx, weight, m):
    n, c = grad_out.shape
    k = idx.shape[1]
    gi = _z((m, c), grad_out.device)
    _sync()
    lib().interpolation_backward_cuda_launcher(n, c, k, _p(grad_out), _p(idx), _p(weight), _p(gi))
    _sync()
    return gi


def subtraction_forward(in1, in2, idx):
    (n, ns), c = idx.shape, in1.shape[1]
    out = _z((n, ns, c), in1.device)
    _sync()
    lib().subtraction_forward_cuda_launcher(n, ns, c, _p(in1), _p(in2), _p(idx), _p(out))
    _sync()
    return out


def subtraction_backward(idx, grad_out, n2=None):
    n, ns, c = grad_out.shape
    g1, g2 = _z((n, c), grad_out.device), _z((n if n2 is None else n2, c), grad_out.device)
    _sync()
    lib().subtraction_backward_cuda_launcher(n, ns, c, _p(idx), _p(grad_out), _p(g1), _p(g2))
    _sync()
    return g1, g2


def aggregation_forward(inp, position, weight, idx):
    (n, ns), c, w_c = idx.shape, inp.shape[1], weight.shape[2]
    out = _z((n, c), inp.device)
    _sync()
    lib().aggregation_forward_cuda_launcher(n, ns, c, w_c, _p(inp), _p(position), _p(weight), _p(idx), _p(out))
    _sync()
    return out


def aggregation_backward(inp, position, weight, idx, grad_out):
    (n, ns), c, w_c = idx.shape, inp.shape[1], weight.shape[2]
    gi, gp, gw = _z(inp.shape, inp.device), _z(position.shape, inp.device), _z(weight.shape, inp.device)
    _sync()
    lib().aggregation_backward_cuda_launcher(n, ns, c, w_c, _p(inp), _p(position), _p(weight), _p(idx), _p(grad_out),
                                             _p(gi), _p(gp), _p(gw))
    _sync()
    return gi, gp, gw


def attention_relation_step_forward(query, key, weight, index_target, index_refer):
    (_, g, c), m = query.shape, index_target.shape[0]
    out = _z((m, g), query.device)
    _sync()
    lib().attention_relation_step_forward_cuda_launcher(m, g, c, _p(query), _p(key), _p(weight), _p(index_target),
                                                        _p(index_refer), _p(out))
    _sync()
    return out


def attention_relation_step_backward(query, key, weight, index_target, index_refer, grad_out):
    (_, g, c), m = query.shape, index_target.shape[0]
    gq, gk, gw = _z(query.shape, query.device), _z(key.shape, query.device), _z(weight.shape, query.device)
    _sync()
    lib().attention_relation_step_backward_cuda_launcher(m, g, c, _p(query), _p(gq), _p(key), _p(gk), _p(weight), _p(gw),
                                                         _p(index_target), _p(index_refer), _p(grad_out))
    _sync()
    return gq, gk, gw


def attention_fusion_step_forward(weight, value, index_target, index_refer):
    (n, g, c), m = value.shape, index_target.shape[0]
    out = _z((n, g, c), value.device)
    _sync()
    lib().attention_fusion_step_forward_cuda_launcher(m, g, c, _p(weight), _p(value), _p(index_target), _p(index_refer), _p(out))
    _sync()
    return out


def attention_fusion_step_backward(weight, value, index_target, index_refer, grad_out):
    (n, g, c), m = value.shape, index_target.shape[0]
    gw, gv = _z(weight.shape, value.device), _z(value.shape, value.device)
    _sync()
    lib().attention_fusion_step_backward_cuda_launcher(m, g, c, _p(weight), _p(gw), _p(value), _p(gv), _p(index_target),
                                                       _p(index_refer), _p(grad_out))
    _sync()
    return gw, gv
