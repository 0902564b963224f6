// batchnorm.cu -- BatchNorm1d + ReLU over token-major rows (R, C) fp32, forward and backward, for the
// per-point MLP of the PointNet backbone (src/models/components/pcd_encoder/pointnet.py:29-55: SubMConv3d(k=1)
// -> BatchNorm1d(eps 1e-3, momentum 0.01) -> ReLU, five times over (sum N, C) = (65 536, 64..512) in cfg-2) and
// the projector of PCDObsEncoder (pcd_obs_encoder.py:100-121).
//
// HBM-bound: the reference composition (and the ATen path this replaces) makes ~26 B/element forward
// (statistics, normalise, ReLU, bf16 cast for the next GEMM: four passes) and ~38 B/element backward; here
//   forward : statistics pass (4 B read) + apply pass (4 B read, 4 + 2 B written: fp32 activation and the bf16
//             operand of the next layer's GEMM)                                   -> 14 B/element
//   backward: reduce pass (8 B read) + apply pass (8 B read, 4 + 2 B written)      -> 22 B/element
// Column sums are kept per thread over a strip of rows, combined in shared memory and leave as ONE fp64 atomic
// per column per CTA (statistics in fp64: E[y^2] - mean^2 is safe there).  Scale / shift / running statistics
// come from pcm_sa_bn_finalize (csrc/sa_fused.cu), shared with the set-abstraction head.
#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;

// Thread (g, t): row group g of `groups`, column quad t of tpr = C / 4.
struct BnMap {
    int tpr, groups, g, t;
    bool active;
    __device__ BnMap(int C) {
        tpr = C >> 2;
        groups = BN_THREADS / tpr;
        g = threadIdx.x / tpr;
        t = threadIdx.x - g * tpr;
        active = g < groups;
    }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// stats[0][c] += sum_r y[r, c]; stats[1][c] += sum_r y[r, c]^2
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const float* __restrict__ y, long R, int C, int rows_per_cta, double* __restrict__ stats) {
    extern __shared__ float part[];  // [2][groups][C]
    pcm_pdl_wait();
    const BnMap m(C);
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (m.active) {
#pragma unroll 4
        for (long r = r0 + m.g; r < r1; r += m.groups) {
            const float4 v = ldg4(y + r * C + m.t * 4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
        }
        float* ps = part + (size_t)m.g * C + m.t * 4;
        float* pq = part + (size_t)(m.groups + m.g) * C + m.t * 4;
        ps[0] = s.x; ps[1] = s.y; ps[2] = s.z; ps[3] = s.w;
        pq[0] = q.x; pq[1] = q.y; pq[2] = q.z; pq[3] = q.w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int g = 0; g < m.groups; ++g) { a += part[(size_t)g * C + c]; b += part[(size_t)(m.groups + g) * C + c]; }
        atomicAdd(stats + c, a);
        atomicAdd(stats + C + c, b);
    }
}

// out = max(a * y + b, 0)  (fp32 and / or bf16)
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_relu_kernel(const float* __restrict__ y, const float* __restrict__ coef, long n4, int C, int relu,
                     float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16) {
    pcm_pdl_wait();
    const int c4 = C >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const float4 v = ldg4(y + i * 4), a = ldg4(coef + c), b = ldg4(coef + C + c);
        float4 o = make_float4(fmaf(a.x, v.x, b.x), fmaf(a.y, v.y, b.y), fmaf(a.z, v.z, b.z), fmaf(a.w, v.w, b.w));
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (out) reinterpret_cast<float4*>(out)[i] = o;
        if (out_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            reinterpret_cast<uint2*>(out_bf16)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
    }
}

// gstats[0][c] += sum_r dz; gstats[1][c] += sum_r dz * xhat, with dz = dout * [a y + b > 0], xhat = (y - mean) invstd
__global__ void __launch_bounds__(BN_THREADS)
bn_relu_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ coef,
                          long R, int C, int relu, int rows_per_cta, double* __restrict__ gstats) {
    extern __shared__ float part[];
    pcm_pdl_wait();
    const BnMap m(C);
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (m.active) {
        const int c = m.t * 4;
        const float4 a = ldg4(coef + c), b = ldg4(coef + C + c), mu = ldg4(coef + 2 * C + c), is = ldg4(coef + 3 * C + c);
#pragma unroll 2
        for (long r = r0 + m.g; r < r1; r += m.groups) {
            const float4 v = ldg4(y + r * C + c);
            float4 d = ldg4(dout + r * C + c);
            if (relu) {
                d.x = fmaf(a.x, v.x, b.x) > 0.f ? d.x : 0.f; d.y = fmaf(a.y, v.y, b.y) > 0.f ? d.y : 0.f;
                d.z = fmaf(a.z, v.z, b.z) > 0.f ? d.z : 0.f; d.w = fmaf(a.w, v.w, b.w) > 0.f ? d.w : 0.f;
            }
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
            q.x += d.x * (v.x - mu.x) * is.x; q.y += d.y * (v.y - mu.y) * is.y;
            q.z += d.z * (v.z - mu.z) * is.z; q.w += d.w * (v.w - mu.w) * is.w;
        }
        float* ps = part + (size_t)m.g * C + c;
        float* pq = part + (size_t)(m.groups + m.g) * C + c;
        ps[0] = s.x; ps[1] = s.y; ps[2] = s.z; ps[3] = s.w;
        pq[0] = q.x; pq[1] = q.y; pq[2] = q.z; pq[3] = q.w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int g = 0; g < m.groups; ++g) { a += part[(size_t)g * C + c]; b += part[(size_t)(m.groups + g) * C + c]; }
        atomicAdd(gstats + c, a);
        atomicAdd(gstats + C + c, b);
    }
}

// training: dy = gamma invstd (dz - mean(dz) - xhat mean(dz xhat));  eval: dy = a dz.
// CTA 0 also accumulates dgamma += sum dz xhat, dbeta += sum dz (unique writer: plain read-modify-write).
__global__ void __launch_bounds__(BN_THREADS)
bn_relu_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ coef,
                         const double* __restrict__ gstats, long R, int C, int relu, int training,
                         float* __restrict__ dy, __nv_bfloat16* __restrict__ dy_bf16, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, const double* __restrict__ n_total_dev) {
    // n_total_dev (SyncBatchNorm): gstats holds the sums over ALL ranks' rows and *n_total_dev their total row count
    pcm_pdl_wait();
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta) dbeta[c] += (float)gstats[c];
            if (dgamma) dgamma[c] += (float)gstats[C + c];
        }
    }
    const int c4 = C >> 2;
    const long n4 = R * c4;
    const float inv_n = n_total_dev ? (float)(1.0 / *n_total_dev) : 1.0f / (float)R;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const float4 v = ldg4(y + i * 4), a = ldg4(coef + c), b = ldg4(coef + C + c);
        float4 d = ldg4(dout + i * 4);
        if (relu) {
            d.x = fmaf(a.x, v.x, b.x) > 0.f ? d.x : 0.f; d.y = fmaf(a.y, v.y, b.y) > 0.f ? d.y : 0.f;
            d.z = fmaf(a.z, v.z, b.z) > 0.f ? d.z : 0.f; d.w = fmaf(a.w, v.w, b.w) > 0.f ? d.w : 0.f;
        }
        float4 o;
        if (training) {
            const float4 mu = ldg4(coef + 2 * C + c), is = ldg4(coef + 3 * C + c);
            const float s0 = (float)gstats[c] * inv_n, s1 = (float)gstats[c + 1] * inv_n, s2 = (float)gstats[c + 2] * inv_n,
                        s3 = (float)gstats[c + 3] * inv_n;
            const float q0 = (float)gstats[C + c] * inv_n, q1 = (float)gstats[C + c + 1] * inv_n,
                        q2 = (float)gstats[C + c + 2] * inv_n, q3 = (float)gstats[C + c + 3] * inv_n;
            o.x = a.x * (d.x - s0 - (v.x - mu.x) * is.x * q0);
            o.y = a.y * (d.y - s1 - (v.y - mu.y) * is.y * q1);
            o.z = a.z * (d.z - s2 - (v.z - mu.z) * is.z * q2);
            o.w = a.w * (d.w - s3 - (v.w - mu.w) * is.w * q3);
        } else {
            o = make_float4(a.x * d.x, a.y * d.y, a.z * d.z, a.w * d.w);
        }
        if (dy) reinterpret_cast<float4*>(dy)[i] = o;
        if (dy_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            reinterpret_cast<uint2*>(dy_bf16)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
    }
}

inline bool bn_shape_ok(int C) { return C > 0 && (C % 4) == 0 && C / 4 <= BN_THREADS; }
inline int bn_rows_per_cta(long R) {
    long rpc = (R + 148L * 4 - 1) / (148L * 4);
    return (int)(rpc < 32 ? 32 : rpc);
}
inline int bn_flat_grid(long n4) {
    const long g = (n4 + BN_THREADS - 1) / BN_THREADS;
    return (int)(g < 1 ? 1 : (g > 148L * 8 ? 148L * 8 : g));
}

}  // namespace

PCM_API int pcm_bn_stats(long long R, int C, const float* y, double* stats, pcm_stream_t stream) {
    if (R <= 0) return PCM_OK;
    if (!y || !stats) return PCM_EINVAL;
    if (!bn_shape_ok(C)) return PCM_EUNSUPPORTED;
    const int rpc = bn_rows_per_cta(R);
    const int grid = (int)((R + rpc - 1) / rpc);
    const size_t smem = (size_t)2 * (BN_THREADS / (C / 4)) * C * sizeof(float);
    cudaError_t e = pcm_launch(bn_stats_kernel, dim3(grid), dim3(BN_THREADS), smem, pcm_cu_stream(stream), y, (long)R, C, rpc, stats);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_bn_apply_relu(long long R, int C, const float* y, const float* coef, int relu, float* out, void* out_bf16,
                              pcm_stream_t stream) {
    if (R <= 0) return PCM_OK;
    if (!y || !coef || (!out && !out_bf16)) return PCM_EINVAL;
    if (!bn_shape_ok(C)) return PCM_EUNSUPPORTED;
    const long n4 = (long)R * (C / 4);
    cudaError_t e = pcm_launch(bn_apply_relu_kernel, dim3(bn_flat_grid(n4)), dim3(BN_THREADS), 0, pcm_cu_stream(stream), y, coef,
                               n4, C, relu, out, reinterpret_cast<__nv_bfloat16*>(out_bf16));
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_bn_relu_bwd(long long R, int C, const float* dout, const float* y, const float* coef, int relu,
                            int training, double* gstats, float* dy, void* dy_bf16, float* dgamma, float* dbeta,
                            pcm_stream_t stream) {
    if (R <= 0) return PCM_OK;
    if (!dout || !y || !coef || !gstats || (!dy && !dy_bf16)) return PCM_EINVAL;
    if (!bn_shape_ok(C)) return PCM_EUNSUPPORTED;
    cudaStream_t st = pcm_cu_stream(stream);
    const int rpc = bn_rows_per_cta(R);
    const int grid = (int)((R + rpc - 1) / rpc);
    const size_t smem = (size_t)2 * (BN_THREADS / (C / 4)) * C * sizeof(float);
    cudaError_t e = pcm_launch(bn_relu_bwd_reduce_kernel, dim3(grid), dim3(BN_THREADS), smem, st, dout, y, coef, (long)R, C, relu,
                               rpc, gstats);
    if (e != cudaSuccess) return (int)e;
    int r = pcm_launch_status();
    if (r) return r;
    const long n4 = (long)R * (C / 4);
    e = pcm_launch(bn_relu_bwd_apply_kernel, dim3(bn_flat_grid(n4)), dim3(BN_THREADS), 0, st, dout, y, coef,
                   (const double*)gstats, (long)R, C, relu, training, dy, reinterpret_cast<__nv_bfloat16*>(dy_bf16), dgamma, dbeta,
                   (const double*)nullptr);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

// The two passes of pcm_bn_relu_bwd as separate entry points, for SyncBatchNorm (configs/trainer/ddp.yaml:9): the caller
// all-reduces gstats (2, C) between them and passes the global row count in device memory (n_total_dev); dgamma / dbeta are
// taken from the LOCAL sums by the caller (torch.nn.SyncBatchNorm semantics: weight gradients stay per rank until DDP
// averages them), so the apply pass is given NULL for them.
PCM_API int pcm_bn_relu_bwd_reduce(long long R, int C, const float* dout, const float* y, const float* coef, int relu,
                                   double* gstats, pcm_stream_t stream) {
    if (R <= 0) return PCM_OK;
    if (!dout || !y || !coef || !gstats) return PCM_EINVAL;
    if (!bn_shape_ok(C)) return PCM_EUNSUPPORTED;
    const int rpc = bn_rows_per_cta(R);
    const int grid = (int)((R + rpc - 1) / rpc);
    const size_t smem = (size_t)2 * (BN_THREADS / (C / 4)) * C * sizeof(float);
    cudaError_t e = pcm_launch(bn_relu_bwd_reduce_kernel, dim3(grid), dim3(BN_THREADS), smem, pcm_cu_stream(stream), dout, y, coef,
                               (long)R, C, relu, rpc, gstats);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_bn_relu_bwd_apply(long long R, int C, const float* dout, const float* y, const float* coef, int relu,
                                  int training, const double* gstats, const double* n_total_dev, float* dy, void* dy_bf16,
                                  float* dgamma, float* dbeta, pcm_stream_t stream) {
    if (R <= 0) return PCM_OK;
    if (!dout || !y || !coef || !gstats || (!dy && !dy_bf16)) return PCM_EINVAL;
    if (!bn_shape_ok(C)) return PCM_EUNSUPPORTED;
    const long n4 = (long)R * (C / 4);
    cudaError_t e = pcm_launch(bn_relu_bwd_apply_kernel, dim3(bn_flat_grid(n4)), dim3(BN_THREADS), 0, pcm_cu_stream(stream), dout, y,
                               coef, gstats, (long)R, C, relu, training, dy, reinterpret_cast<__nv_bfloat16*>(dy_bf16), dgamma, dbeta,
                               n_total_dev);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}
