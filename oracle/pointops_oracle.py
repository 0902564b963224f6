"""numpy/ctypes front-end of oracle/libpointops_oracle.so (oracle/pointops_oracle.c).

TEST INFRASTRUCTURE ONLY.  Function names and argument meaning follow the reference's Python API
(`libs/pointops/functions/*.py`), operating on numpy arrays on the host.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libpointops_oracle.so"


def build() -> Path:
    src = _HERE / "pointops_oracle.c"
    if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "libpointops_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n: int) -> int:
    return lib().oracle_opt_n_threads(int(n))


def farthest_point_sampling(xyz, offset, new_offset):
    """functions/sampling.py:8-23 -> idx (m,) int32."""
    xyz, offset, new_offset = _f(xyz), _i(offset), _i(new_offset)
    b = offset.shape[0]
    sizes = np.diff(np.concatenate([[0], offset]))
    n_max = int(sizes.max())
    idx = np.zeros(int(new_offset[-1]), dtype=np.int32)
    tmp = np.full(xyz.shape[0], 1e10, dtype=np.float32)
    lib().oracle_farthest_point_sampling(b, n_max, _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))
    return idx


def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None, squared=False):
    """functions/query.py:8-23 -> (idx (m,ns) int32, dist (m,ns) f32)."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    xyz, new_xyz, offset, new_offset = _f(xyz), _f(new_xyz), _i(offset), _i(new_offset)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), dtype=np.int32)
    dist2 = np.zeros((m, nsample), dtype=np.float32)
    lib().oracle_knn_query(m, nsample, _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(dist2))
    return idx, (dist2 if squared else np.sqrt(dist2))


def ball_query(nsample, max_radius, min_radius, xyz, offset, new_xyz=None, new_offset=None, squared=False):
    """functions/query.py:72-107."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    xyz, new_xyz, offset, new_offset = _f(xyz), _f(new_xyz), _i(offset), _i(new_offset)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), dtype=np.int32)
    dist2 = np.zeros((m, nsample), dtype=np.float32)
    ovf = ctypes.c_int(0)
    lib().oracle_ball_query(m, nsample, ctypes.c_float(min_radius), ctypes.c_float(max_radius), _p(xyz),
                            _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(dist2), ctypes.byref(ovf))
    return idx, (dist2 if squared else np.sqrt(dist2))


def random_ball_query(nsample, max_radius, min_radius, xyz, offset, new_xyz, new_offset, order, squared=False):
    """functions/query.py:26-69 with the permutation injected."""
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    xyz, new_xyz, offset, new_offset, order = _f(xyz), _f(new_xyz), _i(offset), _i(new_offset), _i(order)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), dtype=np.int32)
    dist2 = np.zeros((m, nsample), dtype=np.float32)
    lib().oracle_random_ball_query(m, nsample, ctypes.c_float(min_radius), ctypes.c_float(max_radius), _p(order),
                                   _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(dist2))
    return idx, (dist2 if squared else np.sqrt(dist2))


def grouping_forward(input, idx):
    input, idx = _f(input), _i(idx)
    m, ns = idx.shape
    c = input.shape[1]
    out = np.zeros((m, ns, c), dtype=np.float32)
    lib().oracle_grouping_forward(m, ns, c, _p(input), _p(idx), _p(out))
    return out


def grouping_backward(grad_output, idx, n):
    grad_output, idx = _f(grad_output), _i(idx)
    m, ns, c = grad_output.shape
    gi = np.zeros((n, c), dtype=np.float32)
    lib().oracle_grouping_backward(m, ns, c, _p(grad_output), _p(idx), _p(gi))
    return gi


def interpolation_forward(input, idx, weight):
    input, idx, weight = _f(input), _i(idx), _f(weight)
    n, k = idx.shape
    c = input.shape[1]
    out = np.zeros((n, c), dtype=np.float32)
    lib().oracle_interpolation_forward(n, c, k, _p(input), _p(idx), _p(weight), _p(out))
    return out


def interpolation_backward(grad_output, idx, weight, m):
    grad_output, idx, weight = _f(grad_output), _i(idx), _f(weight)
    n, c = grad_output.shape
    k = idx.shape[1]
    gi = np.zeros((m, c), dtype=np.float32)
    lib().oracle_interpolation_backward(n, c, k, _p(grad_output), _p(idx), _p(weight), _p(gi))
    return gi


def aggregation_forward(input, position, weight, idx):
    input, position, weight, idx = _f(input), _f(position), _f(weight), _i(idx)
    n, ns, c = position.shape
    w_c = weight.shape[-1]
    out = np.zeros((n, c), dtype=np.float32)
    lib().oracle_aggregation_forward(n, ns, c, w_c, _p(input), _p(position), _p(weight), _p(idx), _p(out))
    return out


def aggregation_backward(input, position, weight, idx, grad_output):
    input, position, weight, idx, grad_output = _f(input), _f(position), _f(weight), _i(idx), _f(grad_output)
    n, ns, c = position.shape
    w_c = weight.shape[-1]
    gi, gp, gw = np.zeros_like(input), np.zeros_like(position), np.zeros_like(weight)
    lib().oracle_aggregation_backward(n, ns, c, w_c, _p(input), _p(position), _p(weight), _p(idx), _p(grad_output),
                                      _p(gi), _p(gp), _p(gw))
    return gi, gp, gw


def subtraction_forward(input1, input2, idx):
    input1, input2, idx = _f(input1), _f(input2), _i(idx)
    n, c = input1.shape
    ns = idx.shape[1]
    out = np.zeros((n, ns, c), dtype=np.float32)
    lib().oracle_subtraction_forward(n, ns, c, _p(input1), _p(input2), _p(idx), _p(out))
    return out


def subtraction_backward(idx, grad_output):
    idx, grad_output = _i(idx), _f(grad_output)
    n, ns, c = grad_output.shape
    g1, g2 = np.zeros((n, c), dtype=np.float32), np.zeros((n, c), dtype=np.float32)
    lib().oracle_subtraction_backward(n, ns, c, _p(idx), _p(grad_output), _p(g1), _p(g2))
    return g1, g2


def attention_relation_step_forward(query, key, weight, index_target, index_refer):
    query, key, weight, it, ir = _f(query), _f(key), _f(weight), _i(index_target), _i(index_refer)
    _, g, c = query.shape
    m = it.shape[0]
    out = np.zeros((m, g), dtype=np.float32)
    lib().oracle_attention_relation_step_forward(m, g, c, _p(query), _p(key), _p(weight), _p(it), _p(ir), _p(out))
    return out


def attention_relation_step_backward(query, key, weight, index_target, index_refer, grad_output):
    query, key, weight, it, ir, go = _f(query), _f(key), _f(weight), _i(index_target), _i(index_refer), _f(grad_output)
    _, g, c = query.shape
    m = it.shape[0]
    gq, gk, gw = np.zeros_like(query), np.zeros_like(key), np.zeros_like(weight)
    lib().oracle_attention_relation_step_backward(m, g, c, _p(query), _p(gq), _p(key), _p(gk), _p(weight), _p(gw),
                                                  _p(it), _p(ir), _p(go))
    return gq, gk, gw


def attention_fusion_step_forward(weight, value, index_target, index_refer):
    weight, value, it, ir = _f(weight), _f(value), _i(index_target), _i(index_refer)
    n, g, c = value.shape
    m = it.shape[0]
    out = np.zeros((n, g, c), dtype=np.float32)
    lib().oracle_attention_fusion_step_forward(m, g, c, _p(weight), _p(value), _p(it), _p(ir), _p(out))
    return out


def attention_fusion_step_backward(weight, value, index_target, index_refer, grad_output):
    weight, value, it, ir, go = _f(weight), _f(value), _i(index_target), _i(index_refer), _f(grad_output)
    n, g, c = value.shape
    m = it.shape[0]
    gw, gv = np.zeros_like(weight), np.zeros_like(value)
    lib().oracle_attention_fusion_step_backward(m, g, c, _p(weight), _p(gw), _p(value), _p(gv), _p(it), _p(ir), _p(go))
    return gw, gv
