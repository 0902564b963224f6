"""Host-side mirror of `libs/pointops/functions/*.py` over the libpcm_b200 C ABI.

Every function keeps the reference's signature and conventions (float32 contiguous CUDA inputs,
int32 outputs, cumulative-end offsets, -1 padding, sqrt'ed distances) -- cited per function.
Differences that are deliberate and invisible to callers:
  * launches go to torch's CURRENT stream (the reference always uses the legacy default stream);
  * outputs are allocated with torch.empty/zeros on the input's device instead of the legacy
    `torch.cuda.IntTensor(...)` constructors;
  * FPS needs ONE host read (largest cloud, total sample count) instead of the reference's
    per-cloud Python loop (functions/sampling.py:14-17); both can be skipped entirely with the
    keyword hints `n_max=` / `m_total=` (used by the training step for sync-free operation).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .._lib import check, current_stream, lib, ptr, require_cuda


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 else t.int()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (reference kernels are fp32-only), got {t.dtype}")
    if not t.is_contiguous():
        raise AssertionError(f"{name} must be contiguous")  # reference asserts, e.g. query.py:16
    return t


# --------------------------------------------------------------------------------------------
# sampling -- functions/sampling.py:6-26
# --------------------------------------------------------------------------------------------
class FarthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, offset, new_offset, n_max=None, m_total=None):
        """xyz (n,3) f32, offset (b), new_offset (b) -> idx (m) int32, GLOBAL row indices."""
        require_cuda(xyz, offset, new_offset)
        xyz = _f32c(xyz, "xyz")
        b = offset.shape[0]
        offset_i, new_offset_i = _i32(offset).contiguous(), _i32(new_offset).contiguous()
        if n_max is None or m_total is None:
            # one device->host read for both numbers (reference: b reads + .item(), sampling.py:14-17)
            sizes = offset_i.clone()
            sizes[1:] -= offset_i[:-1]
            host = torch.stack([sizes.max(), new_offset_i[b - 1]]).tolist()
            n_max = host[0] if n_max is None else n_max
            m_total = host[1] if m_total is None else m_total
        n_max, m_total = int(n_max), int(m_total)
        idx = torch.zeros(m_total, dtype=torch.int32, device=xyz.device)
        tmp = None
        if n_max > 8192:  # only the large-cloud kernel keeps running minima in global memory
            tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
        if m_total > 0 and n_max > 0:
            check(lib.pcm_farthest_point_sampling(b, n_max, ptr(xyz), ptr(offset_i), ptr(new_offset_i),
                                                  ptr(tmp), ptr(idx), current_stream()),
                  "pcm_farthest_point_sampling")
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None, None


def farthest_point_sampling(xyz, offset, new_offset, n_max=None, m_total=None):
    return FarthestPointSampling.apply(xyz, offset, new_offset, n_max, m_total)


# --------------------------------------------------------------------------------------------
# queries -- functions/query.py:6-112
# --------------------------------------------------------------------------------------------
def _query_prologue(xyz, offset, new_xyz, new_offset):
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    require_cuda(xyz, new_xyz, offset, new_offset)
    xyz, new_xyz = _f32c(xyz, "xyz"), _f32c(new_xyz, "new_xyz")
    return xyz, _i32(offset).contiguous(), new_xyz, _i32(new_offset).contiguous()


class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, offset, new_xyz=None, new_offset=None, with_dist=True):
        """-> idx (m, nsample) int32 (-1 pad), dist (m, nsample) f32 = sqrt(d^2) (query.py:8-23)."""
        xyz, offset, new_xyz, new_offset = _query_prologue(xyz, offset, new_xyz, new_offset)
        m = new_xyz.shape[0]
        idx = torch.empty((m, nsample), dtype=torch.int32, device=xyz.device)
        dist2 = torch.empty((m, nsample), dtype=torch.float32, device=xyz.device) if with_dist else None
        check(lib.pcm_knn_query(offset.shape[0], m, nsample, ptr(xyz), ptr(new_xyz), ptr(offset),
                                ptr(new_offset), ptr(idx), ptr(dist2), current_stream()), "pcm_knn_query")
        ctx.mark_non_differentiable(idx)
        if not with_dist:
            return idx, None
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist)
        return idx, dist

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 6


def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None):
    return KNNQuery.apply(nsample, xyz, offset, new_xyz, new_offset, True)


class BallQuery(Function):
    @staticmethod
    def forward(ctx, nsample, max_radius, min_radius, xyz, offset, new_xyz=None, new_offset=None):
        """query.py:72-107 (note the argument order: max_radius BEFORE min_radius)."""
        xyz, offset, new_xyz, new_offset = _query_prologue(xyz, offset, new_xyz, new_offset)
        assert min_radius < max_radius
        m = new_xyz.shape[0]
        idx = torch.empty((m, nsample), dtype=torch.int32, device=xyz.device)
        dist2 = torch.empty((m, nsample), dtype=torch.float32, device=xyz.device)
        check(lib.pcm_ball_query(offset.shape[0], m, nsample, float(min_radius), float(max_radius),
                                 ptr(xyz), ptr(new_xyz), ptr(offset), ptr(new_offset), ptr(idx),
                                 ptr(dist2), current_stream()), "pcm_ball_query")
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 7


ball_query = BallQuery.apply


class RandomBallQuery(Function):
    @staticmethod
    def forward(ctx, nsample, max_radius, min_radius, xyz, offset, new_xyz=None, new_offset=None, order=None):
        """query.py:26-69; `order` (optional) injects the per-cloud permutation for testing."""
        xyz, offset, new_xyz, new_offset = _query_prologue(xyz, offset, new_xyz, new_offset)
        assert min_radius < max_radius
        m = new_xyz.shape[0]
        if order is None:
            # per-cloud random permutation without a Python loop over clouds: sort random keys
            # within each cloud segment (the reference builds one randperm per cloud, :47-53).
            n = xyz.shape[0]
            pos = torch.arange(n, device=xyz.device)
            cloud = torch.searchsorted(offset.long(), pos, right=True)
            keys = cloud.double() + torch.rand(n, device=xyz.device, dtype=torch.float64)
            order = torch.argsort(keys).int()
        order = _i32(order).contiguous()
        idx = torch.empty((m, nsample), dtype=torch.int32, device=xyz.device)
        dist2 = torch.empty((m, nsample), dtype=torch.float32, device=xyz.device)
        check(lib.pcm_random_ball_query(offset.shape[0], m, nsample, float(min_radius), float(max_radius),
                                        ptr(order), ptr(xyz), ptr(new_xyz), ptr(offset), ptr(new_offset),
                                        ptr(idx), ptr(dist2), current_stream()), "pcm_random_ball_query")
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist

    @staticmethod
    def backward(ctx, *grads):
        return (None,) * 8


def random_ball_query(nsample, max_radius, min_radius, xyz, offset, new_xyz=None, new_offset=None, order=None):
    return RandomBallQuery.apply(nsample, max_radius, min_radius, xyz, offset, new_xyz, new_offset, order)


# --------------------------------------------------------------------------------------------
# grouping -- functions/grouping.py:6-62
# --------------------------------------------------------------------------------------------
class Grouping(Function):
    @staticmethod
    def forward(ctx, input, idx):
        """input (n,c) f32, idx (m,nsample) int32 -> (m,nsample,c) (grouping.py:8-21)."""
        require_cuda(input, idx)
        input = _f32c(input, "input")
        assert idx.is_contiguous()
        idx = _i32(idx)
        m, nsample = idx.shape
        n, c = input.shape
        output = torch.empty((m, nsample, c), dtype=torch.float32, device=input.device)
        check(lib.pcm_grouping_forward(m, nsample, c, ptr(input), ptr(idx), ptr(output), current_stream()),
              "pcm_grouping_forward")
        ctx.n = n
        ctx.save_for_backward(idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        m, nsample, c = grad_output.shape
        grad_input = torch.zeros((ctx.n, c), dtype=torch.float32, device=grad_output.device)
        check(lib.pcm_grouping_backward(m, nsample, c, ptr(grad_output), ptr(idx), ptr(grad_input), current_stream()),
              "pcm_grouping_backward")
        return grad_input, None


grouping2 = Grouping.apply


def grouping(idx, feat, xyz, new_xyz=None, with_xyz=False):
    """Differentiable gather with -1 padding -> zero rows (grouping.py:35-59)."""
    if new_xyz is None:
        new_xyz = xyz
    assert xyz.is_contiguous() and feat.is_contiguous()
    m, nsample, c = idx.shape[0], idx.shape[1], feat.shape[1]
    flat = idx.reshape(-1).long()
    valid = (flat >= 0).to(feat.dtype).unsqueeze(1)
    safe = flat.clamp_min(0)
    # zero rows for padded (-1) neighbours; equivalent to the reference's appended zero row
    grouped_feat = (feat[safe, :] * valid).view(m, nsample, c)
    if not with_xyz:
        return grouped_feat
    assert new_xyz.is_contiguous()
    # reference: (xyz_padded[idx] - new_xyz) * sign(idx + 1): padded rows -> exactly 0
    grouped_xyz = (xyz[safe, :].view(m, nsample, 3) - new_xyz.unsqueeze(1)) * valid.view(m, nsample, 1)
    return torch.cat((grouped_xyz, grouped_feat), -1)


# --------------------------------------------------------------------------------------------
# interpolation -- functions/interpolation.py:8-59
# --------------------------------------------------------------------------------------------
def _interp_weights(xyz, new_xyz, offset, new_offset, k):
    idx, dist = knn_query(k, xyz, offset, new_xyz, new_offset)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=1, keepdim=True)
    return idx, dist_recip / norm


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """Pure-torch variant (interpolation.py:8-21): differentiable w.r.t. feat by indexing."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    idx, weight = _interp_weights(xyz, new_xyz, offset, new_offset, k)
    new_feat = torch.zeros((new_xyz.shape[0], feat.shape[1]), dtype=torch.float32, device=feat.device)
    for i in range(k):
        new_feat = new_feat + feat[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return new_feat


class Interpolation(Function):
    @staticmethod
    def forward(ctx, xyz, new_xyz, input, offset, new_offset, k=3):
        """interpolation.py:24-43."""
        require_cuda(xyz, new_xyz, input)
        assert xyz.is_contiguous() and new_xyz.is_contiguous() and input.is_contiguous()
        idx, weight = _interp_weights(xyz, new_xyz, offset, new_offset, k)
        weight = weight.contiguous()
        n, c, m = new_xyz.shape[0], input.shape[1], input.shape[0]
        output = torch.zeros((n, c), dtype=torch.float32, device=input.device)
        check(lib.pcm_interpolation_forward(n, c, k, ptr(input), ptr(idx), ptr(weight), ptr(output), current_stream()),
              "pcm_interpolation_forward")
        ctx.m, ctx.k = m, k
        ctx.save_for_backward(idx, weight)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, c = grad_output.shape
        grad_input = torch.zeros((ctx.m, c), dtype=torch.float32, device=grad_output.device)
        check(lib.pcm_interpolation_backward(n, c, ctx.k, ptr(grad_output), ptr(idx), ptr(weight), ptr(grad_input),
                                             current_stream()), "pcm_interpolation_backward")
        return None, None, grad_input, None, None, None


interpolation2 = Interpolation.apply


# --------------------------------------------------------------------------------------------
# aggregation / subtraction -- functions/aggregation.py, functions/subtraction.py
# --------------------------------------------------------------------------------------------
class Aggregation(Function):
    @staticmethod
    def forward(ctx, input, position, weight, idx):
        """input (n,c), position (n,ns,c), weight (n,ns,c'), idx (n,ns) -> (n,c) (aggregation.py:8-26)."""
        require_cuda(input, position, weight, idx)
        assert input.is_contiguous() and position.is_contiguous() and weight.is_contiguous()
        idx = _i32(idx).contiguous()
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        output = torch.zeros((n, c), dtype=torch.float32, device=input.device)
        check(lib.pcm_aggregation_forward(n, nsample, c, w_c, ptr(input), ptr(position), ptr(weight), ptr(idx),
                                          ptr(output), current_stream()), "pcm_aggregation_forward")
        ctx.save_for_backward(input, position, weight, idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, position, weight, idx = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        grad_input = torch.zeros_like(input)
        grad_position = torch.zeros_like(position)
        grad_weight = torch.zeros_like(weight)
        check(lib.pcm_aggregation_backward(n, nsample, c, w_c, ptr(input), ptr(position), ptr(weight), ptr(idx),
                                           ptr(grad_output), ptr(grad_input), ptr(grad_position), ptr(grad_weight),
                                           current_stream()), "pcm_aggregation_backward")
        return grad_input, grad_position, grad_weight, None


aggregation = Aggregation.apply


class Subtraction(Function):
    @staticmethod
    def forward(ctx, input1, input2, idx):
        """input1 (n,c), input2 (n,c), idx (n,ns) -> (n,ns,c) (subtraction.py:8-21)."""
        require_cuda(input1, input2, idx)
        assert input1.is_contiguous() and input2.is_contiguous()
        idx = _i32(idx).contiguous()
        n, c = input1.shape
        nsample = idx.shape[-1]
        output = torch.empty((n, nsample, c), dtype=torch.float32, device=input1.device)
        check(lib.pcm_subtraction_forward(n, nsample, c, ptr(input1), ptr(input2), ptr(idx), ptr(output),
                                          current_stream()), "pcm_subtraction_forward")
        ctx.save_for_backward(idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, nsample, c = grad_output.shape
        grad_input1 = torch.zeros((n, c), dtype=torch.float32, device=grad_output.device)
        grad_input2 = torch.zeros((n, c), dtype=torch.float32, device=grad_output.device)
        check(lib.pcm_subtraction_backward(n, nsample, c, ptr(idx), ptr(grad_output), ptr(grad_input1),
                                           ptr(grad_input2), current_stream()), "pcm_subtraction_backward")
        return grad_input1, grad_input2, None


subtraction = Subtraction.apply


# --------------------------------------------------------------------------------------------
# scatter attention -- functions/attention.py
# --------------------------------------------------------------------------------------------
class AttentionRelationStep(Function):
    @staticmethod
    def forward(ctx, query, key, weight, index_target, index_refer):
        """query/key (n,g,c), weight (c), index_* (m) -> relation (m,g) (attention.py:13-39)."""
        require_cuda(query, key, weight, index_target, index_refer)
        assert (query.is_contiguous() and key.is_contiguous() and index_target.is_contiguous()
                and index_refer.is_contiguous() and weight.is_contiguous())
        assert index_target.shape[0] == index_refer.shape[0]
        _, g, c = query.shape
        m = index_target.shape[0]
        it, ir = _i32(index_target), _i32(index_refer)
        output = torch.zeros((m, g), dtype=torch.float32, device=query.device)
        check(lib.pcm_attention_relation_step_forward(m, g, c, ptr(query), ptr(key), ptr(weight), ptr(it), ptr(ir),
                                                      ptr(output), current_stream()),
              "pcm_attention_relation_step_forward")
        ctx.save_for_backward(query, key, weight, it, ir)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        query, key, weight, it, ir = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, g, c = query.shape
        m = it.shape[0]
        grad_query = torch.zeros_like(query)
        grad_key = torch.zeros_like(key)
        grad_weight = torch.zeros_like(weight)
        check(lib.pcm_attention_relation_step_backward(m, g, c, ptr(query), ptr(grad_query), ptr(key), ptr(grad_key),
                                                       ptr(weight), ptr(grad_weight), ptr(it), ptr(ir),
                                                       ptr(grad_output), current_stream()),
              "pcm_attention_relation_step_backward")
        # the reference computes grad_weight but returns None for it (attention.py:61): keep that.
        return grad_query, grad_key, None, None, None


class AttentionFusionStep(Function):
    @staticmethod
    def forward(ctx, weight, value, index_target, index_refer):
        """weight (m,g), value (n,g,c), index_* (m) -> (n,g,c) (attention.py:66-92)."""
        require_cuda(weight, value, index_target, index_refer)
        assert (weight.is_contiguous() and value.is_contiguous() and index_target.is_contiguous()
                and index_refer.is_contiguous())
        assert index_target.shape[0] == index_refer.shape[0]
        n, g, c = value.shape
        m = index_refer.shape[0]
        it, ir = _i32(index_target), _i32(index_refer)
        output = torch.zeros((n, g, c), dtype=torch.float32, device=value.device)
        check(lib.pcm_attention_fusion_step_forward(m, g, c, ptr(weight), ptr(value), ptr(it), ptr(ir), ptr(output),
                                                    current_stream()), "pcm_attention_fusion_step_forward")
        ctx.save_for_backward(weight, value, it, ir)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        weight, value, it, ir = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, g, c = value.shape
        m = it.shape[0]
        grad_weight = torch.zeros_like(weight)
        grad_value = torch.zeros_like(value)
        check(lib.pcm_attention_fusion_step_backward(m, g, c, ptr(weight), ptr(grad_weight), ptr(value),
                                                     ptr(grad_value), ptr(it), ptr(ir), ptr(grad_output),
                                                     current_stream()), "pcm_attention_fusion_step_backward")
        return grad_weight, grad_value, None, None


attention_relation_step = AttentionRelationStep.apply
attention_fusion_step = AttentionFusionStep.apply


# --------------------------------------------------------------------------------------------
# helpers -- functions/utils.py
# --------------------------------------------------------------------------------------------
def knn_query_and_group(feat, xyz, offset=None, new_xyz=None, new_offset=None, idx=None, nsample=None,
                        with_xyz=False):
    """utils.py:5-18."""
    if idx is None:
        assert nsample is not None
        idx, _ = KNNQuery.apply(nsample, xyz, offset, new_xyz, new_offset, False)
    return grouping(idx, feat, xyz, new_xyz, with_xyz), idx


def ball_query_and_group(feat, xyz, offset=None, new_xyz=None, new_offset=None, idx=None, max_radio=None,
                         min_radio=0, nsample=None, with_xyz=False):
    """utils.py:21-42."""
    if idx is None:
        assert nsample is not None and offset is not None
        assert max_radio is not None and min_radio is not None
        idx, _ = ball_query(nsample, max_radio, min_radio, xyz, offset, new_xyz, new_offset)
    return grouping(idx, feat, xyz, new_xyz, with_xyz), idx


def query_and_group(nsample, xyz, new_xyz, feat, idx, offset, new_offset, dilation=0, with_feat=True,
                    with_xyz=True):
    """utils.py:45-99 (dilated kNN grouping; no -1 masking, like the reference)."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        num_samples_total = 1 + (nsample - 1) * (dilation + 1)
        idx_no_dilation, _ = knn_query(num_samples_total, xyz, offset, new_xyz, new_offset)
        ends = offset.tolist()
        starts = [0] + ends[:-1]
        new_ends = new_offset.tolist()
        new_starts = [0] + new_ends[:-1]
        parts = []
        for s, e, ns, ne in zip(starts, ends, new_starts, new_ends):
            soft = (e - s - 1) / (nsample - 1) - 1 if e - s < num_samples_total else dilation
            cols = [int((soft + 1) * j) for j in range(nsample)]
            parts.append(idx_no_dilation[ns:ne, cols])
        idx = torch.cat(parts, dim=0)
    if not with_feat:
        return idx
    m, c = new_xyz.shape[0], feat.shape[1]
    flat = idx.reshape(-1).long()
    grouped_xyz = xyz[flat, :].view(m, nsample, 3) - new_xyz.unsqueeze(1)
    grouped_feat = feat[flat, :].view(m, nsample, c)
    if with_xyz:
        return torch.cat((grouped_xyz, grouped_feat), -1), idx
    return grouped_feat, idx


def offset2batch(offset):
    """utils.py:102-117, without the per-cloud Python loop / host sync."""
    offset = offset.long()
    n = offset[-1]
    sizes = torch.diff(offset, prepend=offset.new_zeros(1))
    return torch.repeat_interleave(torch.arange(offset.shape[0], device=offset.device), sizes)


def batch2offset(batch):
    """utils.py:120-121."""
    return torch.cumsum(batch.bincount(), dim=0).int()
