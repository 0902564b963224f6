#!/usr/bin/env python
"""Target for compute-sanitizer (memcheck / racecheck / synccheck): one small invocation of every kernel family of
libpcm_b200.so -- the smoke path (FPS, kNN, fused set abstraction, tcgen05 GEMMs, fused attention fwd/bwd, LayerNorm,
BatchNorm, token / head kernels, optimizer, Diffusion-Policy U-Net kernels) plus the grid-sample and sparse-convolution
kernels -- at sizes that finish in minutes under the tools' 10-100x slowdown.

    compute-sanitizer --tool memcheck  --log-file profiles/rN_sanitizer_memcheck.log  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck --log-file profiles/rN_sanitizer_racecheck.log python tools/sanitize_target.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

G.smoke()
from pointcloudmatters_b200.data_gpu import collate_raw_clouds  # noqa: E402
from pointcloudmatters_b200.spunet import SpUNet  # noqa: E402

rng = np.random.default_rng(0)
clouds = [(rng.normal(0, 0.03, (1500, 3)).astype(np.float32), rng.integers(0, 256, (1500, 3)).astype(np.float32)) for _ in range(3)]
out = collate_raw_clouds(clouds, "cuda", grid_size=0.01, mode="train", seed=1)
torch.manual_seed(0)
net = SpUNet(6, num_classes=16, base_channels=16, channels=(16, 32, 32, 32), layers=(1, 1, 1, 1)).cuda().train()
y = net(dict(grid_coord=out["grid_coord"], feat=out["feat"], offset=out["offset"]))
y.pow(2).sum().backward()
torch.cuda.synchronize()
# warp-per-query kNN: a lattice cloud (distance ties -> the exact heap-replay phase) and nsample > 31 (thread-per-query kernel)
from pointcloudmatters_b200 import pointops as P  # noqa: E402

lat = torch.from_numpy((rng.integers(0, 4, (700, 3)) / 4.0).astype(np.float32)).cuda()
off = torch.tensor([300, 700], dtype=torch.int32, device="cuda")
for k in (16, 33):
    ki, kd = P.knn_query(k, lat, off, lat[::3].contiguous(), torch.tensor([100, 234], dtype=torch.int32, device="cuda"))
# raw-frame filter (both dataset families) feeding the voxel-grid stage
from pointcloudmatters_b200.data_gpu import filter_frames_maniskill2, filter_frames_rlbench, grid_sample_collate  # noqa: E402

fx = torch.from_numpy(rng.uniform(-1, 1, (2, 16384, 4)).astype(np.float32)).cuda()
fc = torch.from_numpy(rng.integers(0, 256, (2, 16384, 3)).astype(np.uint8)).cuda()
c3, col3, off3 = filter_frames_maniskill2(fx, fc, crop=np.array([[3, 9], [0, 15]], np.int32))
grid_sample_collate(c3, col3, off3, grid_size=0.05)
rp = torch.from_numpy(rng.uniform(-1, 2, (2, 2, 64, 64, 3)).astype(np.float32)).cuda()
filter_frames_rlbench(rp, rp * 100, torch.from_numpy(rng.integers(0, 250, (2, 2, 64, 64)).astype(np.float32)).cuda())
torch.cuda.synchronize()
print("sanitize target OK:", tuple(y.shape), tuple(ki.shape), tuple(c3.shape))
