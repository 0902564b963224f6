import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pointcloudmatters_b200.data_gpu import filter_frames_maniskill2, grid_sample_collate
rng = np.random.default_rng(0)
b, P = 64, 16384
xyzw = rng.uniform(-0.5, 0.5, (b, P, 4)).astype(np.float32); xyzw[..., 3] = (rng.uniform(0, 1, (b, P)) > 0.3); xyzw[..., 2] = np.abs(xyzw[..., 2])
rgb = rng.integers(0, 256, (b, P, 3)).astype(np.uint8)
X, C = torch.from_numpy(xyzw).cuda(), torch.from_numpy(rgb).cuda()
for _ in range(3):
    c, col, off = filter_frames_maniskill2(X, C); out = grid_sample_collate(c, col, off, grid_size=0.005)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    c, col, off = filter_frames_maniskill2(X, C)
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(20):
    out = grid_sample_collate(c, col, off, grid_size=0.005)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"64 frames x 16384 px: filter {1e3*(t1-t0)/20:.3f} ms ({c.shape[0]} survivors), grid sample + collate {1e3*(t2-t1)/20:.3f} ms ({out['coord'].shape[0]} voxels)")
