"""GPU data path: voxel-grid subsampling + collation of a whole batch of raw clouds on the device.

Replaces the per-sample CPU transforms the reference runs in 16 loader workers (`configs/data/
maniskill2_act_pcd_dataset.yaml:15-34`): `GridSamplePCD(grid_size, hash_type="fnv", return_grid_coord=True,
keys=[coord, color])` (src/data/components/transformpcd.py:684-793) -> `NormalizeColorPCD` -> `CollectPCD(feat_keys=
[color, coord])` -> `pcd_collate_fn` (src/utils/sparse_tensor_utils.py:65-82), and returns the `pcds` dict of the batch
contract together with the `n_max` hint that keeps the training step free of device->host reads.

One device->host read per batch (the B voxel counts: the packed tensors have data-dependent sizes); it belongs to the
loader stage, not to the training step.  CUDA only; kernels in csrc/grid_sample.cu through the C ABI.
"""
from __future__ import annotations

import torch

from ._lib import PcmError, check, current_stream, lib, ptr

_INT_MAX = 2 ** 31 - 1


def grid_sample_collate(coord, color, offset, grid_size=0.005, mode="test", seed=0, color_scale=127.5, color_shift=1.0,
                        append_coord=True, prio=None, f32_div=False, return_index=False):
    """coord (N, 3) f32, color (N, C) f32 (raw 0..255 values), offset (B) int64 cumulative ends -- all on the GPU.
    mode "test": the first point of every voxel survives (the reference's test-mode part 0); "train": a uniformly random
    member (per-point random priorities from `seed`); `prio` (N) uint32-valued int64 tensor overrides both.
    Returns {coord, grid_coord, feat, offset, n_max[, index]}."""
    if not coord.is_cuda:
        raise PcmError("grid_sample_collate runs on CUDA tensors only (no CPU fallback)")
    n, b = coord.shape[0], offset.shape[0]
    dev = coord.device
    coord = coord.contiguous().float()
    color = color.contiguous().float() if color is not None else None
    off64 = offset.to(torch.int64).contiguous()
    gs = (grid_size,) * 3 if isinstance(grid_size, (int, float)) else tuple(grid_size)
    if prio is None and mode == "train":
        g = torch.Generator(device=dev).manual_seed(int(seed))
        prio = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, device=dev, dtype=torch.int64)
    prio32 = prio.to(torch.int64).to(torch.int32).contiguous() if prio is not None else None  # reinterpreted as uint32
    grid = torch.empty((n, 3), dtype=torch.int32, device=dev)
    gmin = torch.full((b, 3), _INT_MAX, dtype=torch.int32, device=dev)
    tkey = torch.full((2 * n,), -1, dtype=torch.int64, device=dev)
    tbest = torch.full((2 * n,), -1, dtype=torch.int64, device=dev)
    skey = torch.empty((2 * n,), dtype=torch.int64, device=dev)
    sval = torch.empty((2 * n,), dtype=torch.int32, device=dev)
    idx_raw = torch.empty((n,), dtype=torch.int64, device=dev)
    grid_raw = torch.empty((n, 3), dtype=torch.int64, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    st = current_stream()
    check(lib.pcm_grid_sample_select(b, n, ptr(coord), ptr(off64), float(gs[0]), float(gs[1]), float(gs[2]), int(f32_div),
                                     ptr(prio32), ptr(grid), ptr(gmin), ptr(tkey), ptr(tbest), ptr(skey), ptr(sval), ptr(idx_raw),
                                     ptr(grid_raw), ptr(counts), st), "pcm_grid_sample_select")
    new_off = torch.cumsum(counts.to(torch.int64), 0)
    host = torch.stack([new_off[-1], counts.max().to(torch.int64)]).cpu()  # the one device->host read of the batch
    m, n_max = int(host[0]), int(host[1])
    fc = color.shape[1] if color is not None else 0
    oc = fc + (3 if append_coord else 0)
    coord_out = torch.empty((m, 3), dtype=torch.float32, device=dev)
    grid_out = torch.empty((m, 3), dtype=torch.int64, device=dev)
    feat_out = torch.empty((m, oc), dtype=torch.float32, device=dev)
    index_out = torch.empty((m,), dtype=torch.int64, device=dev) if return_index else None
    check(lib.pcm_grid_sample_gather(b, m, ptr(off64), ptr(new_off), ptr(idx_raw), ptr(grid_raw), ptr(coord), ptr(color), fc,
                                     float(color_scale), float(color_shift), int(append_coord), ptr(coord_out), ptr(grid_out),
                                     ptr(feat_out), ptr(index_out), st), "pcm_grid_sample_gather")
    out = {"coord": coord_out, "grid_coord": grid_out, "feat": feat_out, "offset": new_off, "n_max": n_max}
    if return_index:
        out["index"] = index_out
    return out


def collate_raw_clouds(samples, device, **kw):
    """`pcd_collate_fn` for RAW clouds: `samples` = list of (coord (n_i, 3), color (n_i, C)) host arrays / tensors, one
    per cloud.  Packs them into pinned staging buffers, copies once, and runs `grid_sample_collate` on the device."""
    import numpy as np

    sizes = [int(c.shape[0]) for c, _ in samples]
    total = sum(sizes)
    fc = samples[0][1].shape[1]
    coord = torch.empty((total, 3), dtype=torch.float32).pin_memory()
    color = torch.empty((total, fc), dtype=torch.float32).pin_memory()
    pos = 0
    for (c, f), s in zip(samples, sizes):
        coord[pos:pos + s] = torch.as_tensor(np.asarray(c), dtype=torch.float32)
        color[pos:pos + s] = torch.as_tensor(np.asarray(f), dtype=torch.float32)
        pos += s
    offset = torch.tensor(sizes, dtype=torch.int64).cumsum(0)
    return grid_sample_collate(coord.to(device, non_blocking=True), color.to(device, non_blocking=True), offset.to(device), **kw)


def _filter_frames(mode, xyz, color, *, include_ground=False, bounds=None, crop=None, cam_hw=(128, 128), crop_size=112, seg=None,
                   invalid=()):
    """Shared driver of the two frame filters (csrc/frame_filter.cu): flag + count per 1024-point chunk, exclusive scan,
    ordered scatter.  Returns (coord (N, 3) f32, color (N, C) f32, offset (B) int64) on the device of `xyz`; ONE
    device->host read (the total survivor count sizes the outputs) -- it belongs to the loader stage."""
    import ctypes

    if not xyz.is_cuda:
        raise PcmError("the frame filters run on CUDA tensors only (no CPU fallback)")
    b, P, stride = xyz.shape
    dev = xyz.device
    xyz = xyz.contiguous().float()
    color_u8 = color.dtype == torch.uint8
    color = color.contiguous() if color_u8 else color.contiguous().float()
    cc = color.shape[-1]
    chunks = (P + 1023) // 1024
    counts = torch.empty((b, chunks), dtype=torch.int32, device=dev)
    bnd = (ctypes.c_double * 6)(*[float(v) for v in (bounds if bounds is not None else (0,) * 6)])
    bnd_p = ctypes.addressof(bnd) if bounds is not None else None
    crop_t = None
    if crop is not None:
        crop_t = torch.as_tensor(crop, dtype=torch.int32).reshape(b, 2).to(dev).contiguous()
    seg_t = seg.contiguous().float() if seg is not None else None
    inv_t = torch.tensor([float(v) for v in invalid], dtype=torch.float32, device=dev) if (seg is not None and len(invalid)) else None
    st = current_stream()
    check(lib.pcm_frame_filter_count(b, P, mode, ptr(xyz), stride, int(include_ground), bnd_p, ptr(crop_t), cam_hw[0], cam_hw[1],
                                     crop_size, ptr(counts), st), "pcm_frame_filter_count")
    incl = torch.cumsum(counts.view(-1).to(torch.int64), 0)
    base = (incl - counts.view(-1)).contiguous()
    offset = incl.view(b, chunks)[:, -1].contiguous()
    n = int(offset[-1])  # the one device->host read
    out_xyz = torch.empty((n, 3), dtype=torch.float32, device=dev)
    out_ch = cc + (1 if seg is not None else 0)
    out_color = torch.empty((n, out_ch), dtype=torch.float32, device=dev)
    check(lib.pcm_frame_filter_scatter(b, P, mode, ptr(xyz), stride, int(include_ground), bnd_p, ptr(crop_t), cam_hw[0], cam_hw[1],
                                       crop_size, ptr(color), int(color_u8), cc, ptr(seg_t), ptr(inv_t),
                                       0 if inv_t is None else inv_t.numel(), ptr(base), ptr(out_xyz), ptr(out_color), st),
          "pcm_frame_filter_scatter")
    return out_xyz, out_color, offset


def filter_frames_maniskill2(xyzw, rgb, include_ground=False, crop=None, cam_hw=(128, 128), crop_size=112):
    """ManiSkill2 frames of a batch -> packed clouds (maniskill2_single_task_pcd_act.py:196-224): xyzw (B, P, 4) f32 and rgb
    (B, P, 3) uint8 / f32 of the selected cameras (P = cams * 128 * 128, camera-major), keep w > 0 and z > 0.005 (or
    x > -0.8 with `include_ground`); `crop` = per-sample (first row, first column) of the `rand_crop` window or None.
    Returns (coord (N, 3), color (N, 3) raw 0..255 values, offset (B)): the inputs of `grid_sample_collate`."""
    return _filter_frames(0, xyzw, rgb, include_ground=include_ground, crop=crop, cam_hw=cam_hw, crop_size=crop_size)


RLBENCH_SCENE_BOUNDS = (-0.3, -0.5, 0.6, 0.7, 0.5, 1.6)  # src/data/components/rlbench/constants.py:1


def filter_frames_rlbench(point_maps, rgbs, masks=None, bounds=RLBENCH_SCENE_BOUNDS, invalid_mask_values=(201, 204, 208, 246)):
    """RLBench multi-view fusion + scene-bounds crop of a batch (rlbench_single_task_act.py:266-295): point_maps / rgbs
    (B, cams, h, w, 3) (or already flattened (B, P, 3)), masks (B, cams, h, w) instance ids or None.  The cameras are
    concatenated camera-major, points strictly inside `bounds` survive; with `masks` the colours get a fourth {0, 1}
    channel (invalid ids -> 0, other ids > 0 -> 1).  Returns (coord (N, 3), color (N, 3 | 4), offset (B))."""
    b = point_maps.shape[0]
    xyz = point_maps.reshape(b, -1, 3)
    col = rgbs.reshape(b, -1, 3)
    seg = masks.reshape(b, -1) if masks is not None else None
    return _filter_frames(1, xyz, col, bounds=bounds, seg=seg, invalid=invalid_mask_values)
