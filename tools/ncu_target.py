"""Tiny ncu target: a few launches of one kernel family at a BASELINE cfg-2 shape.
usage: python tools/ncu_target.py gemm|gemm_bf16out|fps|knn|sa|flash|hbm
(`hbm`: the HBM-bound row kernels at the main-encoder size -- add+dropout+LayerNorm fwd / bwd on 32 960 x 512,
BatchNorm+ReLU fwd / bwd on 65 536 x 512, colsum)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if which == "attn_s":
    pass
elif which == "flash":
    from pointcloudmatters_b200 import kernels as K

    B, nh, L, S = 64, 8, 515, 515
    Z, E = B * nh, nh * 64
    q = torch.randn(Z * L, 64, device="cuda").bfloat16()
    k = torch.randn(Z * S, 64, device="cuda").bfloat16()
    v = torch.randn(Z * S, 64, device="cuda").bfloat16()
    do = torch.randn(Z * L, 64, device="cuda").bfloat16()
    sb = torch.tensor([1234567], dtype=torch.int64, device="cuda")
    buf = torch.empty(L * B, 3 * E, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        O, lse = K.flash_attn_fwd(q, k, v, B, nh, L, S, None, 0.125, 0.1, sb, 77)
        K.flash_attn_bwd(q, k, v, O, do, lse, B, nh, L, S, None, 0.125, 0.1, sb, 77, buf[:, :E], buf[:, E:2 * E], buf[:, 2 * E:])
elif which == "hbm":
    from pointcloudmatters_b200 import kernels as K
    from pointcloudmatters_b200 import functional as PF

    rows, C = 32960, 512
    sb = torch.tensor([1234567], dtype=torch.int64, device="cuda")
    x, res = torch.randn(rows, C, device="cuda"), torch.randn(rows, C, device="cuda")
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    pos = torch.randn(rows // 64, C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    for _ in range(2):
        y, yb, h, mean, rstd, ypb = K.add_dropout_ln_fwd(x, res, g, b, 1e-5, 0.1, sb, 7, want_bf16=True, pos=pos, pos_row_div=64)
        K.add_dropout_ln_bwd(y, h, mean, rstd, g, 0.1, sb, 7, True, dg, db, True, dy_b=res)
        K.colsum(yb, dg)
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
    z = torch.randn(65536, C, device="cuda", requires_grad=True)
    for _ in range(2):
        o = PF.batchnorm_relu(z, bn)
        o.backward(torch.ones_like(o))
elif which == "gemm_grouped":
    # the decoder bucket's queue of one training step: 7 layers x (self dWo, dW_qk, dW_v, cross dWo, dW_q, FFN dW1, dW2)
    from pointcloudmatters_b200.kernels import gemm_dw_grouped

    rows = 6400
    mk = lambda r, c: torch.randn(r, c, device="cuda").bfloat16()
    probs = []
    for _ in range(7):
        for (m, n) in ((512, 512), (1024, 512), (512, 512), (512, 512), (512, 512), (32, 512), (512, 32)):
            probs.append((mk(rows, m), mk(rows, n), torch.zeros(m, n, device="cuda")))
    for _ in range(3):
        gemm_dw_grouped(probs)
elif which == "gemm_inproj":
    # fused Q|K|V in-projection of a main-encoder layer: two A operands, parted head-split bf16 output
    from pointcloudmatters_b200 import kernels as K

    L, B, E, nh = 515, 64, 512, 8
    Z = B * nh
    xqk, xv = torch.randn(L * B, E, device="cuda").bfloat16(), torch.randn(L * B, E, device="cuda").bfloat16()
    w, bias = torch.randn(3 * E, E, device="cuda").bfloat16(), torch.randn(3 * E, device="cuda")
    qkv = torch.empty(3, Z * L, 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        K.gemm_ex(L * B, 3 * E, E, 1, xqk, False, 0, w, False, 0, qkv, c_mode=1, hs=(B, nh, L), ldc=64, bias=bias, a2=xv,
                  a2_from_col=2 * E, hs_parts=(E, Z * L * 64))
elif which.startswith("gemm"):
    from pointcloudmatters_b200.kernels import gemm_bf16

    M, N, K = 32960, 1024, 512
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if which == "gemm_bf16out" else torch.float32)
    for _ in range(4):
        gemm_bf16(a, b, out=out)
else:
    import numpy as np

    from pointcloudmatters_b200 import pointops as P
    from tests._data import clouds

    xyz, off, noff = clouds(64, 1024, 512, seed=1)
    t_xyz, t_off, t_noff = [torch.from_numpy(x).cuda() for x in (xyz, off, noff)]
    for _ in range(4):
        fps = P.farthest_point_sampling(t_xyz, t_off, t_noff, n_max=1024, m_total=64 * 512)
        q = t_xyz[fps.long()].contiguous()
        P.knn_query(16, t_xyz, t_off, q, t_noff)
torch.cuda.synchronize()
if which == "attn_s":
    from pointcloudmatters_b200.kernels import gemm_ex
    Z, L, S = 512, 515, 515
    Qh = torch.randn(Z * L, 64, device="cuda").bfloat16()
    Kh = torch.randn(Z * S, 64, device="cuda").bfloat16()
    Sbuf = torch.empty(Z * 576, 576, device="cuda")
    for _ in range(4):
        gemm_ex(L, S, 64, Z, Qh, False, L, Kh, False, S, Sbuf, c_mode=0, c_batch_rows=576, ldc=576)
    torch.cuda.synchronize()
