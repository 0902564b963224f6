import sys
sys.path.insert(0, "/root/repo")
import torch
from pointcloudmatters_b200.kernels import gemm_bf16
M, N, K = 6400, 512, 512
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
o32 = torch.empty(M, N, device="cuda")
o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bt = torch.randn(K, N, device="cuda").bfloat16()
for _ in range(3):
    gemm_bf16(a, b, out=o32, bias=bias)
    gemm_bf16(a, b, out=o16, bias=bias)
    gemm_bf16(a, bt, b_mn=True, out=o32)
torch.cuda.synchronize()
