// sa_fused.cu -- fused set-abstraction head for sm_100a.
//
// Replaces the chain in ACTPCD.pcd_sampling (reference src/models/components/act/act.py:446-460):
//     grouping(idx, feat, xyz, new_xyz, with_xyz=True)      -> (m, k, 3+C)    1.08 GB at cfg-2
//     Linear(3+C -> H, bias=False)                          -> (m, k, H)      276 GFLOP fp32 SGEMM
//     transpose/contiguous, BatchNorm1d(H) (batch stats), ReLU, MaxPool1d(k)   3 more 1 GB passes
// by an algebraically exact reformulation that never materialises an (m, k, .) tensor:
//     y[m,j,c] = W[c,:] . [xyz[i]-q_m, feat[i]] = Pf[i,c] + Wx[c,:] . (xyz[i] - q_m),   i = idx[m,j]
//     Pf = feat Wf^T  -- ONE (n x C) x (C x H) GEMM over SOURCE points on tcgen05 (36 GFLOP);
//     BatchNorm is a per-channel monotone affine map, so
//     out[m,c] = ReLU(a_c * (a_c >= 0 ? max_j y : min_j y) + b_c).
// Forward = one gather pass (pcm_sa_gather_stats: per (m,c) max / min / arg + per-channel sum,
// sum of squares and sum of y*dxyz in fp64) + BN finalize + an (m x H) elementwise pass.
// Backward is exact for training-mode BN: dy[m,j,c] = a_c*(delta_{j=j*} dz[m,c] - A_c - xhat*B_c)
// splits into a SPARSE part (one scatter per (m,c)) and a DENSE part that is affine in y and
// therefore collapses per source point: dPf[i,c] = cnt_i*(alpha_c + beta_c*Pf[i,c]) +
// beta_c * Wx[c,:].(cnt_i*xyz_i - SQ_i).  All kernels are HBM/L2-bound gathers: one thread owns 4
// channels (128-bit loads of a Pf row), a CTA owns a strip of queries, per-channel partial sums
// live in registers and reach global memory as one fp64 atomic per channel per CTA.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int SA_STAT_ROWS = 5;  // s1, s2, sum y*dx, sum y*dy, sum y*dz   (forward)
                                 // dbeta, dgamma, sparse dWx[0..2]        (backward)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ---- forward gather ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sa_gather_stats_kernel(
    const float* __restrict__ Pf, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
    const int* __restrict__ idx, const float* __restrict__ W, int ldw, int m, int k, int H,
    float* __restrict__ ymax, float* __restrict__ ymin, unsigned char* __restrict__ jmax,
    unsigned char* __restrict__ jmin, double* __restrict__ stats) {
    const int c0 = threadIdx.x * 4;
    const bool act = c0 < H;
    float wx[4][3];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int d = 0; d < 3; ++d) wx[v][d] = act ? __ldg(W + (size_t)(c0 + v) * ldw + d) : 0.f;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0}, sy[4] = {0, 0, 0, 0}, sz[4] = {0, 0, 0, 0};
    for (int q = blockIdx.x; q < m; q += gridDim.x) {
        const float qx = __ldg(new_xyz + (size_t)q * 3 + 0), qy = __ldg(new_xyz + (size_t)q * 3 + 1), qz = __ldg(new_xyz + (size_t)q * 3 + 2);
        float mx[4], mn[4];
        int amx[4] = {0, 0, 0, 0}, amn[4] = {0, 0, 0, 0};
#pragma unroll
        for (int v = 0; v < 4; ++v) { mx[v] = -INFINITY; mn[v] = INFINITY; }
        const int* qi = idx + (size_t)q * k;
        for (int j = 0; j < k; ++j) {
            const int i = __ldg(qi + j);
            float dx = 0.f, dy = 0.f, dz = 0.f;
            float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i >= 0) {  // -1 = padding: the reference groups an all-zero row (y = 0, still counted by BN)
                dx = __ldg(xyz + (size_t)i * 3 + 0) - qx;
                dy = __ldg(xyz + (size_t)i * 3 + 1) - qy;
                dz = __ldg(xyz + (size_t)i * 3 + 2) - qz;
                if (act) pf = ld4(Pf + (size_t)i * H + c0);
            }
            const float pv[4] = {pf.x, pf.y, pf.z, pf.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float y = fmaf(wx[v][2], dz, fmaf(wx[v][1], dy, fmaf(wx[v][0], dx, pv[v])));
                if (y > mx[v]) { mx[v] = y; amx[v] = j; }
                if (y < mn[v]) { mn[v] = y; amn[v] = j; }
                s1[v] += y;
                s2[v] = fmaf(y, y, s2[v]);
                sx[v] = fmaf(y, dx, sx[v]);
                sy[v] = fmaf(y, dy, sy[v]);
                sz[v] = fmaf(y, dz, sz[v]);
            }
        }
        if (act) {
            *reinterpret_cast<float4*>(ymax + (size_t)q * H + c0) = make_float4(mx[0], mx[1], mx[2], mx[3]);
            *reinterpret_cast<float4*>(ymin + (size_t)q * H + c0) = make_float4(mn[0], mn[1], mn[2], mn[3]);
            *reinterpret_cast<uchar4*>(jmax + (size_t)q * H + c0) = make_uchar4(amx[0], amx[1], amx[2], amx[3]);
            *reinterpret_cast<uchar4*>(jmin + (size_t)q * H + c0) = make_uchar4(amn[0], amn[1], amn[2], amn[3]);
        }
    }
    if (act) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            atomicAdd(stats + 0 * H + c0 + v, (double)s1[v]);
            atomicAdd(stats + 1 * H + c0 + v, (double)s2[v]);
            atomicAdd(stats + 2 * H + c0 + v, (double)sx[v]);
            atomicAdd(stats + 3 * H + c0 + v, (double)sy[v]);
            atomicAdd(stats + 4 * H + c0 + v, (double)sz[v]);
        }
    }
}

// one edge (neighbour j) of a thread's four channels: y = Pf + Wx . d (on s y in single-extreme mode), extreme tracking,
// the five running sums.  Same operation order as the scalar generic kernel.
template <bool ONE>
__device__ __forceinline__ void sa_edge(const float4 pf, float dx, float dy, float dz, int j, const float2 (&wxp)[2][3],
                                        const float2 (&sg2)[2], float (&mx)[4], int (&amx)[4], float (&mn)[4], int (&amn)[4],
                                        float2 (&s1)[2], float2 (&s2)[2], float2 (&sx)[2], float2 (&sy)[2], float2 (&sz)[2]) {
    const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float2 pv = h == 0 ? make_float2(pf.x, pf.y) : make_float2(pf.z, pf.w);
        if (ONE) pv = pcm_fmul2(pv, sg2[h]);
        const float2 y = pcm_ffma2(wxp[h][2], dz2, pcm_ffma2(wxp[h][1], dy2, pcm_ffma2(wxp[h][0], dx2, pv)));
        if (y.x > mx[2 * h]) { mx[2 * h] = y.x; amx[2 * h] = j; }
        if (y.y > mx[2 * h + 1]) { mx[2 * h + 1] = y.y; amx[2 * h + 1] = j; }
        if (!ONE) {
            if (y.x < mn[2 * h]) { mn[2 * h] = y.x; amn[2 * h] = j; }
            if (y.y < mn[2 * h + 1]) { mn[2 * h + 1] = y.y; amn[2 * h + 1] = j; }
        }
        s1[h] = pcm_fadd2(s1[h], y);
        s2[h] = pcm_ffma2(y, y, s2[h]);
        sx[h] = pcm_ffma2(y, dx2, sx[h]);
        sy[h] = pcm_ffma2(y, dy2, sy[h]);
        sz[h] = pcm_ffma2(y, dz2, sz[h]);
    }
}

// Same pass for a compile-time neighbour count K <= 32 (the reference configuration has nsample = 16).  In the generic
// kernel above every thread walks the k neighbours through a dependent chain (index load -> three coordinate loads and the
// 16-byte Pf gather), one neighbour at a time.  Here lane j of each warp fetches neighbour j's index and offset vector
// once per query -- K independent loads -- and the loop, fully unrolled, receives them by shuffle: the K Pf gathers of a
// thread are independent instructions the compiler can issue ahead of the arithmetic.  Identical arithmetic, identical
// results.
template <int K, bool ONE>
__global__ void __launch_bounds__(256, 3) sa_gather_stats_k_kernel(
    const float* __restrict__ Pf, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
    const int* __restrict__ idx, const float* __restrict__ W, int ldw, const float* __restrict__ sel_gamma, int m, int H,
    float* __restrict__ ymax, float* __restrict__ ymin, unsigned char* __restrict__ jmax,
    unsigned char* __restrict__ jmin, double* __restrict__ stats) {
    const int c0 = threadIdx.x * 4, lane = threadIdx.x & 31;
    const bool act = c0 < H;
    float wx[4][3];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int d = 0; d < 3; ++d) wx[v][d] = act ? __ldg(W + (size_t)(c0 + v) * ldw + d) : 0.f;
    // ONE: only the extreme BatchNorm will select is tracked -- sign(a_c) = sign(gamma_c) is known before the statistics --
    // by running the whole pass on s y (s = -1 where gamma < 0): max(s y) is max or min of y, the sums are mirrored exactly
    float sg[4] = {1.f, 1.f, 1.f, 1.f};
    if (ONE) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            sg[v] = (act && __ldg(sel_gamma + c0 + v) < 0.f) ? -1.f : 1.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) wx[v][d] *= sg[v];
        }
    }
    float2 wxp[2][3], sg2[2], s1p[2], s2p[2], sxp[2], syp[2], szp[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        sg2[h] = make_float2(sg[2 * h], sg[2 * h + 1]);
        s1p[h] = s2p[h] = sxp[h] = syp[h] = szp[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 3; ++d) wxp[h][d] = make_float2(wx[2 * h][d], wx[2 * h + 1][d]);
    }
    for (int q = blockIdx.x; q < m; q += gridDim.x) {
        const float qx = __ldg(new_xyz + (size_t)q * 3 + 0), qy = __ldg(new_xyz + (size_t)q * 3 + 1), qz = __ldg(new_xyz + (size_t)q * 3 + 2);
        int ni = -1;
        float ndx = 0.f, ndy = 0.f, ndz = 0.f;
        if (lane < K) {
            ni = __ldg(idx + (size_t)q * K + lane);
            if (ni >= 0) {  // -1 = padding: the reference groups an all-zero row (y = 0, still counted by BN)
                ndx = __ldg(xyz + (size_t)ni * 3 + 0) - qx;
                ndy = __ldg(xyz + (size_t)ni * 3 + 1) - qy;
                ndz = __ldg(xyz + (size_t)ni * 3 + 2) - qz;
            }
        }
        float mx[4], mn[4];
        int amx[4] = {0, 0, 0, 0}, amn[4] = {0, 0, 0, 0};
#pragma unroll
        for (int v = 0; v < 4; ++v) { mx[v] = -INFINITY; mn[v] = INFINITY; }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int i = __shfl_sync(PCM_FULL_MASK, ni, j);
            const float dx = __shfl_sync(PCM_FULL_MASK, ndx, j), dy = __shfl_sync(PCM_FULL_MASK, ndy, j),
                        dz = __shfl_sync(PCM_FULL_MASK, ndz, j);
            float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i >= 0 && act) pf = ld4(Pf + (size_t)i * H + c0);
            sa_edge<ONE>(pf, dx, dy, dz, j, wxp, sg2, mx, amx, mn, amn, s1p, s2p, sxp, syp, szp);
        }
        if (act) {
            *reinterpret_cast<float4*>(ymax + (size_t)q * H + c0) = make_float4(mx[0] * sg[0], mx[1] * sg[1], mx[2] * sg[2], mx[3] * sg[3]);
            *reinterpret_cast<uchar4*>(jmax + (size_t)q * H + c0) = make_uchar4(amx[0], amx[1], amx[2], amx[3]);
            if (!ONE) {
                *reinterpret_cast<float4*>(ymin + (size_t)q * H + c0) = make_float4(mn[0], mn[1], mn[2], mn[3]);
                *reinterpret_cast<uchar4*>(jmin + (size_t)q * H + c0) = make_uchar4(amn[0], amn[1], amn[2], amn[3]);
            }
        }
    }
    float s1[4] = {s1p[0].x, s1p[0].y, s1p[1].x, s1p[1].y}, s2[4] = {s2p[0].x, s2p[0].y, s2p[1].x, s2p[1].y};
    float sx[4] = {sxp[0].x, sxp[0].y, sxp[1].x, sxp[1].y}, sy[4] = {syp[0].x, syp[0].y, syp[1].x, syp[1].y};
    float sz[4] = {szp[0].x, szp[0].y, szp[1].x, szp[1].y};
    if (ONE) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { s1[v] *= sg[v]; sx[v] *= sg[v]; sy[v] *= sg[v]; sz[v] *= sg[v]; }
    }
    if (act) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            atomicAdd(stats + 0 * H + c0 + v, (double)s1[v]);
            atomicAdd(stats + 1 * H + c0 + v, (double)s2[v]);
            atomicAdd(stats + 2 * H + c0 + v, (double)sx[v]);
            atomicAdd(stats + 3 * H + c0 + v, (double)sy[v]);
            atomicAdd(stats + 4 * H + c0 + v, (double)sz[v]);
        }
    }
}

// ---- BatchNorm finalize: coef = [a, b, mean, invstd] (4 x H), running-stat update -------------
__global__ void sa_bn_finalize_kernel(const double* __restrict__ stats, double n_rows, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, float momentum, int training,
                                      float* __restrict__ running_mean, float* __restrict__ running_var,
                                      float* __restrict__ coef, int H, const double* __restrict__ n_rows_dev) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    if (n_rows_dev) n_rows = *n_rows_dev;  // SyncBatchNorm: global row count, all-reduced together with the statistics
    double mean, var;
    if (training) {
        mean = stats[c] / n_rows;
        var = stats[H + c] / n_rows - mean * mean;
        if (var < 0) var = 0;
        if (running_mean) {
            const double unbiased = n_rows > 1 ? var * n_rows / (n_rows - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double a = (double)gamma[c] * invstd;
    coef[0 * H + c] = (float)a;
    coef[1 * H + c] = (float)((double)beta[c] - mean * a);
    coef[2 * H + c] = (float)mean;
    coef[3 * H + c] = (float)invstd;
}

// ---- out = ReLU(a * ext + b); jsel = arg of the selected extreme -------------------------------
__global__ void __launch_bounds__(256) sa_output_kernel(const float* __restrict__ ymax, const float* __restrict__ ymin,
                                                        const unsigned char* __restrict__ jmax,
                                                        const unsigned char* __restrict__ jmin,
                                                        const float* __restrict__ coef, long total, int H,
                                                        float* __restrict__ out, unsigned char* __restrict__ jsel) {
    for (long e = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < total; e += (long)gridDim.x * blockDim.x * 4) {
        const int c = (int)(e % H);
        const float4 a = ld4(coef + c), b = ld4(coef + H + c);
        // ymin == NULL: single-extreme intermediates (pcm_sa_gather_sel*): ymax / jmax already hold the selected extreme
        const float4 hi = ld4(ymax + e), lo = ymin ? ld4(ymin + e) : hi;
        const uchar4 jh = *reinterpret_cast<const uchar4*>(jmax + e), jl = jmin ? *reinterpret_cast<const uchar4*>(jmin + e) : jh;
        float4 o;
        uchar4 js;
        o.x = fmaxf(fmaf(a.x, a.x >= 0.f ? hi.x : lo.x, b.x), 0.f); js.x = a.x >= 0.f ? jh.x : jl.x;
        o.y = fmaxf(fmaf(a.y, a.y >= 0.f ? hi.y : lo.y, b.y), 0.f); js.y = a.y >= 0.f ? jh.y : jl.y;
        o.z = fmaxf(fmaf(a.z, a.z >= 0.f ? hi.z : lo.z, b.z), 0.f); js.z = a.z >= 0.f ? jh.z : jl.z;
        o.w = fmaxf(fmaf(a.w, a.w >= 0.f ? hi.w : lo.w, b.w), 0.f); js.w = a.w >= 0.f ? jh.w : jl.w;
        *reinterpret_cast<float4*>(out + e) = o;
        *reinterpret_cast<uchar4*>(jsel + e) = js;
    }
}

// ---- token-layout variant of sa_output_kernel -------------------------------------------------
// Query q = b * per + mi of cloud b is written to row (head + mi) * batch + b of the transformer's seq-first
// (S, B, H) token tensor (transformer.py:75-92 builds it with flatten / permute / cat passes), together with the
// bf16 operand copies of the first encoder layer's projections: bf16(out) and bf16(out + pos).
__global__ void __launch_bounds__(256) sa_output_tokens_kernel(
    const float* __restrict__ ymax, const float* __restrict__ ymin, const unsigned char* __restrict__ jmax,
    const unsigned char* __restrict__ jmin, const float* __restrict__ coef, long total, int H, int per, int batch, int head,
    const float* __restrict__ pos, float* __restrict__ out, __nv_bfloat16* __restrict__ out_b,
    __nv_bfloat16* __restrict__ out_pb, unsigned char* __restrict__ jsel) {
    for (long e = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < total; e += (long)gridDim.x * blockDim.x * 4) {
        const long q = e / H;
        const int c = (int)(e - q * H);
        const long row = (long)(head + (int)(q % per)) * batch + q / per;
        const long d = row * H + c;
        const float4 a = ld4(coef + c), b = ld4(coef + H + c);
        // ymin == NULL: single-extreme intermediates (pcm_sa_gather_sel*): ymax / jmax already hold the selected extreme
        const float4 hi = ld4(ymax + e), lo = ymin ? ld4(ymin + e) : hi;
        const uchar4 jh = *reinterpret_cast<const uchar4*>(jmax + e), jl = jmin ? *reinterpret_cast<const uchar4*>(jmin + e) : jh;
        float4 o;
        uchar4 js;
        o.x = fmaxf(fmaf(a.x, a.x >= 0.f ? hi.x : lo.x, b.x), 0.f); js.x = a.x >= 0.f ? jh.x : jl.x;
        o.y = fmaxf(fmaf(a.y, a.y >= 0.f ? hi.y : lo.y, b.y), 0.f); js.y = a.y >= 0.f ? jh.y : jl.y;
        o.z = fmaxf(fmaf(a.z, a.z >= 0.f ? hi.z : lo.z, b.z), 0.f); js.z = a.z >= 0.f ? jh.z : jl.z;
        o.w = fmaxf(fmaf(a.w, a.w >= 0.f ? hi.w : lo.w, b.w), 0.f); js.w = a.w >= 0.f ? jh.w : jl.w;
        *reinterpret_cast<float4*>(out + d) = o;
        *reinterpret_cast<uchar4*>(jsel + e) = js;
        if (out_b) {
            __nv_bfloat162 l2 = __floats2bfloat162_rn(o.x, o.y), h2 = __floats2bfloat162_rn(o.z, o.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&l2);
            pk.y = *reinterpret_cast<uint32_t*>(&h2);
            *reinterpret_cast<uint2*>(out_b + d) = pk;
        }
        if (out_pb) {
            const float4 p4 = ld4(pos + d);
            __nv_bfloat162 l2 = __floats2bfloat162_rn(o.x + p4.x, o.y + p4.y), h2 = __floats2bfloat162_rn(o.z + p4.z, o.w + p4.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&l2);
            pk.y = *reinterpret_cast<uint32_t*>(&h2);
            *reinterpret_cast<uint2*>(out_pb + d) = pk;
        }
    }
}

// ---- backward pass 1: per-channel reductions + sparse scatter ----------------------------------
// gstats rows: 0 dbeta = sum dz, 1 dgamma = sum dz*xhat_sel, 2..4 sum a*dz*dxyz_sel (sparse dWx)
__global__ void __launch_bounds__(256) sa_bwd_scatter_kernel(
    const float* __restrict__ dout, const float* __restrict__ out, const unsigned char* __restrict__ jsel,
    const int* __restrict__ idx, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
    const float* __restrict__ coef, int m, int k, int H, float* __restrict__ dPf, double* __restrict__ gstats,
    int per, int batch, int head, const float* __restrict__ dout2) {
    // per > 0: dout / out (and the optional second gradient dout2, summed on load) are in the token layout of
    // sa_output_tokens_kernel; per == 0: query-major rows.
    const int c0 = threadIdx.x * 4;
    const bool act = c0 < H;
    float a[4], b[4], mean[4], invstd[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        a[v] = act ? coef[0 * H + c0 + v] : 0.f;
        b[v] = act ? coef[1 * H + c0 + v] : 0.f;
        mean[v] = act ? coef[2 * H + c0 + v] : 0.f;
        invstd[v] = act ? coef[3 * H + c0 + v] : 0.f;
    }
    float g0[4] = {0, 0, 0, 0}, g1[4] = {0, 0, 0, 0}, gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0}, gz[4] = {0, 0, 0, 0};
    for (int q = blockIdx.x; q < m; q += gridDim.x) {
        if (!act) continue;
        const float qx = __ldg(new_xyz + (size_t)q * 3 + 0), qy = __ldg(new_xyz + (size_t)q * 3 + 1), qz = __ldg(new_xyz + (size_t)q * 3 + 2);
        const size_t row = per > 0 ? (size_t)(head + q % per) * batch + q / per : (size_t)q;
        float4 d4 = ld4(dout + row * H + c0);
        const float4 o4 = ld4(out + row * H + c0);
        if (dout2) {
            const float4 e4 = ld4(dout2 + row * H + c0);
            d4.x += e4.x; d4.y += e4.y; d4.z += e4.z; d4.w += e4.w;
        }
        const uchar4 j4 = *reinterpret_cast<const uchar4*>(jsel + (size_t)q * H + c0);
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
        const int jv[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (ov[v] > 0.f && dv[v] != 0.f) {
                const float dz = dv[v];
                const float ysel = a[v] != 0.f ? (ov[v] - b[v]) / a[v] : mean[v];
                g0[v] += dz;
                g1[v] = fmaf(dz, (ysel - mean[v]) * invstd[v], g1[v]);
                const int i = __ldg(idx + (size_t)q * k + jv[v]);
                if (i >= 0) {
                    const float adz = a[v] * dz;
                    atomicAdd(dPf + (size_t)i * H + c0 + v, adz);
                    gx[v] = fmaf(adz, __ldg(xyz + (size_t)i * 3 + 0) - qx, gx[v]);
                    gy[v] = fmaf(adz, __ldg(xyz + (size_t)i * 3 + 1) - qy, gy[v]);
                    gz[v] = fmaf(adz, __ldg(xyz + (size_t)i * 3 + 2) - qz, gz[v]);
                }
            }
        }
    }
    if (act) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            atomicAdd(gstats + 0 * H + c0 + v, (double)g0[v]);
            atomicAdd(gstats + 1 * H + c0 + v, (double)g1[v]);
            atomicAdd(gstats + 2 * H + c0 + v, (double)gx[v]);
            atomicAdd(gstats + 3 * H + c0 + v, (double)gy[v]);
            atomicAdd(gstats + 4 * H + c0 + v, (double)gz[v]);
        }
    }
}

// ---- edge statistics: cnt[i], SQ[i] = sum of q over incoming edges, total sum of dxyz ---------
__global__ void __launch_bounds__(256) sa_edge_stats_kernel(const int* __restrict__ idx, const float* __restrict__ xyz,
                                                            const float* __restrict__ new_xyz, long edges, int k,
                                                            float* __restrict__ cnt, float* __restrict__ sq,
                                                            double* __restrict__ sdtot) {
    float tx = 0.f, ty = 0.f, tz = 0.f;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < edges; e += (long)gridDim.x * blockDim.x) {
        const int i = __ldg(idx + e);
        if (i < 0) continue;
        const long q = e / k;
        const float qx = __ldg(new_xyz + q * 3 + 0), qy = __ldg(new_xyz + q * 3 + 1), qz = __ldg(new_xyz + q * 3 + 2);
        atomicAdd(cnt + i, 1.0f);
        atomicAdd(sq + (size_t)i * 3 + 0, qx);
        atomicAdd(sq + (size_t)i * 3 + 1, qy);
        atomicAdd(sq + (size_t)i * 3 + 2, qz);
        tx += __ldg(xyz + (size_t)i * 3 + 0) - qx;
        ty += __ldg(xyz + (size_t)i * 3 + 1) - qy;
        tz += __ldg(xyz + (size_t)i * 3 + 2) - qz;
    }
    for (int o = 16; o > 0; o >>= 1) {
        tx += __shfl_xor_sync(PCM_FULL_MASK, tx, o);
        ty += __shfl_xor_sync(PCM_FULL_MASK, ty, o);
        tz += __shfl_xor_sync(PCM_FULL_MASK, tz, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sdtot + 0, (double)tx);
        atomicAdd(sdtot + 1, (double)ty);
        atomicAdd(sdtot + 2, (double)tz);
    }
}

// ---- backward coefficients: alpha, beta' (2 x H) + the small parameter gradients --------------
// dW[:, 0:3] (pitch ldw) += sparse + alpha*SDtot + beta'*SYD ; dgamma, dbeta.
__global__ void sa_bwd_coef_kernel(const double* __restrict__ gstats, const double* __restrict__ fstats,
                                   const double* __restrict__ sdtot, const float* __restrict__ coef, double n_rows,
                                   int training, int H, float* __restrict__ ab, float* __restrict__ dW, int ldw,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta, const double* __restrict__ n_rows_dev) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    if (n_rows_dev) n_rows = *n_rows_dev;
    const double a = coef[c], mean = coef[2 * H + c], invstd = coef[3 * H + c];
    const double dbe = gstats[c], dga = gstats[H + c];
    double alpha = 0.0, betap = 0.0;
    if (training) {
        const double A = dbe / n_rows, B = dga / n_rows;
        betap = -a * B * invstd;
        alpha = -a * A - betap * mean;
    }
    ab[c] = (float)alpha;
    ab[H + c] = (float)betap;
    if (dgamma) dgamma[c] = (float)dga;
    if (dbeta) dbeta[c] = (float)dbe;
    if (dW) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
            dW[(size_t)c * ldw + d] = (float)(gstats[(2 + d) * H + c] + alpha * sdtot[d] + betap * fstats[(2 + d) * H + c]);
    }
}

// ---- backward dense part over source points; emits the bf16 operand of the two dPf GEMMs ------
__global__ void __launch_bounds__(256) sa_bwd_dense_kernel(const float* __restrict__ Pf, const float* __restrict__ xyz,
                                                           const float* __restrict__ cnt, const float* __restrict__ sq,
                                                           const float* __restrict__ W, int ldw,
                                                           const float* __restrict__ ab, int n, int H,
                                                           const float* __restrict__ dPf,
                                                           __nv_bfloat16* __restrict__ dPf_bf16) {
    const int c0 = threadIdx.x * 4;
    if (c0 >= H) return;
    float wx[4][3], al[4], be[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int d = 0; d < 3; ++d) wx[v][d] = __ldg(W + (size_t)(c0 + v) * ldw + d);
        al[v] = ab[c0 + v];
        be[v] = ab[H + c0 + v];
    }
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const float cn = __ldg(cnt + i);
        const float ex = cn * __ldg(xyz + (size_t)i * 3 + 0) - __ldg(sq + (size_t)i * 3 + 0);
        const float ey = cn * __ldg(xyz + (size_t)i * 3 + 1) - __ldg(sq + (size_t)i * 3 + 1);
        const float ez = cn * __ldg(xyz + (size_t)i * 3 + 2) - __ldg(sq + (size_t)i * 3 + 2);
        const float4 pf = ld4(Pf + (size_t)i * H + c0), sp = ld4(dPf + (size_t)i * H + c0);
        const float pv[4] = {pf.x, pf.y, pf.z, pf.w}, sv[4] = {sp.x, sp.y, sp.z, sp.w};
        float r[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float wd = fmaf(wx[v][2], ez, fmaf(wx[v][1], ey, wx[v][0] * ex));
            r[v] = sv[v] + cn * fmaf(be[v], pv[v], al[v]) + be[v] * wd;
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dPf_bf16 + (size_t)i * H + c0) = pk;
    }
}

// ---- cloud-slice variant: the Pf rows of ONE cloud, restricted to a slice of CS channels, live in shared memory ----
// The gather kernels above re-read every Pf row about M k / N = 8 times (1.07 GB of L2 -> SM traffic at cfg-2).  Here a CTA owns (cloud, CS-channel slice): it stages the slice of the cloud's Pf rows once
// (n_b x CS fp32 -- 64 KB at 1024 points x 16 channels; Pf is then read from global memory exactly once in total), walks the
// cloud's queries with CS / 4 threads per query, and gathers from shared memory.  Neighbour indices and offset vectors of a
// query are fetched once by the query's threads (K / LPR neighbours each) and exchanged by shuffle.  Same per-element
// arithmetic as the generic kernel: identical ymax / ymin / arg; the per-channel sums differ in summation order only (fp32
// per thread and warp, fp64 across warps).  cfg-2: 284 us (generic kernel 433 us, its shuffle-unrolled form 307 us); 256 us
// tracking the selected extreme only, 240 us with packed FFMA2 / FADD2 arithmetic (100 registers: 2 CTAs per SM; capping the
// registers for a third CTA spills and costs more than it gains, the 128-thread kernel above is capped at 80 instead).
// per-channel partial sums of a cloud-slice CTA: lanes with equal lane % LPR hold the same four channels
template <int LPR, int CSV>
__device__ __forceinline__ void sa_slice_reduce_impl(float (&r0)[4], float (&r1)[4], float (&r2)[4], float (&r3)[4],
                                                     float (&r4)[4], double (*red)[CSV], int cl, int lane) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            r0[v] += __shfl_xor_sync(PCM_FULL_MASK, r0[v], o);
            r1[v] += __shfl_xor_sync(PCM_FULL_MASK, r1[v], o);
            r2[v] += __shfl_xor_sync(PCM_FULL_MASK, r2[v], o);
            r3[v] += __shfl_xor_sync(PCM_FULL_MASK, r3[v], o);
            r4[v] += __shfl_xor_sync(PCM_FULL_MASK, r4[v], o);
        }
        if (lane < LPR) {
            atomicAdd(&red[0][cl * 4 + v], (double)r0[v]);
            atomicAdd(&red[1][cl * 4 + v], (double)r1[v]);
            atomicAdd(&red[2][cl * 4 + v], (double)r2[v]);
            atomicAdd(&red[3][cl * 4 + v], (double)r3[v]);
            atomicAdd(&red[4][cl * 4 + v], (double)r4[v]);
        }
    }
}
template <int LPR>
__device__ __forceinline__ void sa_slice_reduce(float (&r0)[4], float (&r1)[4], float (&r2)[4], float (&r3)[4], float (&r4)[4],
                                                double (*red)[LPR * 4], int cl, int lane) {
    sa_slice_reduce_impl<LPR, LPR * 4>(r0, r1, r2, r3, r4, red, cl, lane);
}

template <int CS, int K, bool ONE>
__global__ void __launch_bounds__(256) sa_gather_stats_cloud_kernel(
    const float* __restrict__ Pf, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
    const int* __restrict__ idx, const int* __restrict__ offset, const int* __restrict__ new_offset,
    const float* __restrict__ W, int ldw, const float* __restrict__ sel_gamma, int H, int n_cap, float* __restrict__ ymax, float* __restrict__ ymin,
    unsigned char* __restrict__ jmax, unsigned char* __restrict__ jmin, double* __restrict__ stats) {
    extern __shared__ __align__(16) float4 slice4[];  // [n_b][LPR]
    __shared__ double red[SA_STAT_ROWS][CS];
    constexpr int LPR = CS / 4;     // threads (float4 lanes) per row / per query
    constexpr int QS = 256 / LPR;   // queries in flight per CTA
    constexpr int QPW = 32 / LPR;   // queries per warp
    constexpr int NPL = K / LPR;    // neighbours fetched per thread
    static_assert(K % LPR == 0 && NPL >= 1, "K must be a multiple of the threads per query");
    const int cloud = blockIdx.y, c_base = blockIdx.x * CS;
    const int s_n = cloud ? __ldg(offset + cloud - 1) : 0;
    const int n_b = min(__ldg(offset + cloud) - s_n, n_cap);
    const int s_m = cloud ? __ldg(new_offset + cloud - 1) : 0, e_m = __ldg(new_offset + cloud);
    for (int t = threadIdx.x; t < n_b * LPR; t += 256) {
        const int r = t / LPR, l = t - r * LPR;
        slice4[t] = ld4(Pf + (size_t)(s_n + r) * H + c_base + 4 * l);
    }
    for (int t = threadIdx.x; t < SA_STAT_ROWS * CS; t += 256) (&red[0][0])[t] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cl = lane % LPR, grp = lane - cl;  // grp = first lane of this query's thread group
    const int c0 = c_base + cl * 4;
    float wx[4][3];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int d = 0; d < 3; ++d) wx[v][d] = __ldg(W + (size_t)(c0 + v) * ldw + d);
    // ONE: only the extreme BatchNorm will select is tracked -- sign(a_c) = sign(gamma_c) is known before the statistics --
    // by running the whole pass on s y (s = -1 where gamma < 0): max(s y) is max or min of y, the sums are mirrored exactly
    float sg[4] = {1.f, 1.f, 1.f, 1.f};
    if (ONE) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            sg[v] = __ldg(sel_gamma + c0 + v) < 0.f ? -1.f : 1.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) wx[v][d] *= sg[v];
        }
    }
    float2 wxp[2][3], sg2[2], s1p[2], s2p[2], sxp[2], syp[2], szp[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        sg2[h] = make_float2(sg[2 * h], sg[2 * h + 1]);
        s1p[h] = s2p[h] = sxp[h] = syp[h] = szp[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int d = 0; d < 3; ++d) wxp[h][d] = make_float2(wx[2 * h][d], wx[2 * h + 1][d]);
    }
    for (int qw = s_m + warp * QPW; qw < e_m; qw += QS) {  // warp-uniform trip count (the shuffles need every lane)
        const int q = qw + lane / LPR;
        const bool valid = q < e_m;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (valid) { qx = __ldg(new_xyz + (size_t)q * 3 + 0); qy = __ldg(new_xyz + (size_t)q * 3 + 1); qz = __ldg(new_xyz + (size_t)q * 3 + 2); }
        int ni[NPL];
        float ndx[NPL], ndy[NPL], ndz[NPL];
#pragma unroll
        for (int u = 0; u < NPL; ++u) {  // this thread fetches neighbours u * LPR + cl of its query
            ni[u] = -1; ndx[u] = ndy[u] = ndz[u] = 0.f;
            if (valid) {
                const int i = __ldg(idx + (size_t)q * K + u * LPR + cl);
                if (i >= 0) {  // -1 = padding: the reference groups an all-zero row (y = 0, still counted by BN)
                    ndx[u] = __ldg(xyz + (size_t)i * 3 + 0) - qx;
                    ndy[u] = __ldg(xyz + (size_t)i * 3 + 1) - qy;
                    ndz[u] = __ldg(xyz + (size_t)i * 3 + 2) - qz;
                    const int il = i - s_n;
                    ni[u] = (il >= 0 && il < n_b) ? il : -2;  // -2: outside the staged rows (undersized hint): treated as padding
                }
            }
        }
        float mx[4], mn[4];
        int amx[4] = {0, 0, 0, 0}, amn[4] = {0, 0, 0, 0};
#pragma unroll
        for (int v = 0; v < 4; ++v) { mx[v] = -INFINITY; mn[v] = INFINITY; }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int src = grp + (j % LPR), u = j / LPR;
            const int i = __shfl_sync(PCM_FULL_MASK, ni[u], src);
            const float dx = __shfl_sync(PCM_FULL_MASK, ndx[u], src), dy = __shfl_sync(PCM_FULL_MASK, ndy[u], src),
                        dz = __shfl_sync(PCM_FULL_MASK, ndz[u], src);
            float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i >= 0) pf = slice4[i * LPR + cl];
            sa_edge<ONE>(pf, dx, dy, dz, j, wxp, sg2, mx, amx, mn, amn, s1p, s2p, sxp, syp, szp);
        }
        if (valid) {
            *reinterpret_cast<float4*>(ymax + (size_t)q * H + c0) = make_float4(mx[0] * sg[0], mx[1] * sg[1], mx[2] * sg[2], mx[3] * sg[3]);
            *reinterpret_cast<uchar4*>(jmax + (size_t)q * H + c0) = make_uchar4(amx[0], amx[1], amx[2], amx[3]);
            if (!ONE) {
                *reinterpret_cast<float4*>(ymin + (size_t)q * H + c0) = make_float4(mn[0], mn[1], mn[2], mn[3]);
                *reinterpret_cast<uchar4*>(jmin + (size_t)q * H + c0) = make_uchar4(amn[0], amn[1], amn[2], amn[3]);
            }
        }
    }
    float s1[4] = {s1p[0].x, s1p[0].y, s1p[1].x, s1p[1].y}, s2[4] = {s2p[0].x, s2p[0].y, s2p[1].x, s2p[1].y};
    float sx[4] = {sxp[0].x, sxp[0].y, sxp[1].x, sxp[1].y}, sy[4] = {syp[0].x, syp[0].y, syp[1].x, syp[1].y};
    float sz[4] = {szp[0].x, szp[0].y, szp[1].x, szp[1].y};
    if (ONE) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { s1[v] *= sg[v]; sx[v] *= sg[v]; sy[v] *= sg[v]; sz[v] *= sg[v]; }
    }
    // threads of a warp that own the same channels (lane % LPR equal) are summed by shuffle; one shared atomic per warp
    sa_slice_reduce<LPR>(s1, s2, sx, sy, sz, red, cl, lane);
    __syncthreads();
    for (int t = threadIdx.x; t < SA_STAT_ROWS * CS; t += 256) {
        const int r = t / CS, c = t - r * CS;
        atomicAdd(stats + (size_t)r * H + c_base + c, red[r][c]);
    }
}

// slice width for a cloud bound of n_max points: the widest of 32 / 16 / 8 / 4 channels that divides H and keeps three CTAs
// per SM (<= 72 KB of shared memory each); failing that, the widest that fits one CTA; 0 = use the generic kernels
inline int sa_slice_width(int n_max, int H) {
    const int cand[4] = {32, 16, 8, 4};
    for (int c : cand)
        if (H % c == 0 && (size_t)n_max * c * 4 <= 72 * 1024) return c;
    for (int c : cand)
        if (H % c == 0 && (size_t)n_max * c * 4 <= 200 * 1024) return c;
    return 0;
}
template <typename Kern>
inline cudaError_t sa_allow_smem(Kern kern, size_t smem) {
    return smem > 48 * 1024 ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
}

inline int sa_threads(int H) { return ((H / 4 + 31) / 32) * 32; }
inline int sa_grid(long work) { long g = 148L * 8; return (int)(work < g ? (work > 0 ? work : 1) : g); }

}  // namespace

PCM_API int pcm_sa_gather_stats(int m, int k, int H, const float* Pf, const float* xyz, const float* new_xyz,
                                const int* idx, const float* W, int ldw, float* ymax, float* ymin,
                                unsigned char* jmax, unsigned char* jmin, double* stats, pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (!Pf || !xyz || !new_xyz || !idx || !W || !ymax || !ymin || !jmax || !jmin || !stats) return PCM_EINVAL;
    if (H % 4 || H > 4096 || k > 255 || k <= 0) return PCM_EUNSUPPORTED;
    static const bool generic_only = [] { const char* e = getenv("PCM_SA_GENERIC"); return e && e[0] == '1'; }();  // A/B timing
    if (k == 16 && !generic_only)
        sa_gather_stats_k_kernel<16, false><<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(Pf, xyz, new_xyz, idx, W, ldw, nullptr,
                                                                                                    m, H, ymax, ymin, jmax, jmin, stats);
    else if (k == 32 && !generic_only)
        sa_gather_stats_k_kernel<32, false><<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(Pf, xyz, new_xyz, idx, W, ldw, nullptr,
                                                                                                    m, H, ymax, ymin, jmax, jmin, stats);
    else
        sa_gather_stats_kernel<<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(Pf, xyz, new_xyz, idx, W, ldw, m, k, H, ymax,
                                                                                       ymin, jmax, jmin, stats);
    return pcm_launch_status();
}

PCM_API int pcm_sa_bn_finalize(int H, const double* stats, double n_rows, const float* gamma, const float* beta,
                               float eps, float momentum, int training, float* running_mean, float* running_var,
                               float* coef, pcm_stream_t stream) {
    return pcm_sa_bn_finalize_ex(H, stats, n_rows, nullptr, gamma, beta, eps, momentum, training, running_mean, running_var, coef,
                                 stream);
}

// n_rows_dev != NULL: the row count is read from device memory (SyncBatchNorm: statistics and count all-reduced over ranks)
PCM_API int pcm_sa_bn_finalize_ex(int H, const double* stats, double n_rows, const double* n_rows_dev, const float* gamma,
                                  const float* beta, float eps, float momentum, int training, float* running_mean,
                                  float* running_var, float* coef, pcm_stream_t stream) {
    if (H <= 0) return PCM_OK;
    if (!stats || !gamma || !beta || !coef || (!training && (!running_mean || !running_var))) return PCM_EINVAL;
    sa_bn_finalize_kernel<<<pcm_divup(H, 128), 128, 0, pcm_cu_stream(stream)>>>(stats, n_rows, gamma, beta, eps, momentum, training,
                                                                               running_mean, running_var, coef, H, n_rows_dev);
    return pcm_launch_status();
}

PCM_API int pcm_sa_output(int m, int H, const float* ymax, const float* ymin, const unsigned char* jmax,
                          const unsigned char* jmin, const float* coef, float* out, unsigned char* jsel,
                          pcm_stream_t stream) {
    const long total = (long)m * H;
    if (total <= 0) return PCM_OK;
    if (!ymax || !jmax || !coef || !out || !jsel || (!ymin != !jmin)) return PCM_EINVAL;
    if (H % 4) return PCM_EUNSUPPORTED;
    sa_output_kernel<<<sa_grid((total / 4 + 255) / 256), 256, 0, pcm_cu_stream(stream)>>>(ymax, ymin, jmax, jmin, coef, total, H, out, jsel);
    return pcm_launch_status();
}

PCM_API int pcm_sa_bwd_scatter(int m, int k, int H, const float* dout, const float* out, const unsigned char* jsel,
                               const int* idx, const float* xyz, const float* new_xyz, const float* coef, float* dPf,
                               double* gstats, pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (!dout || !out || !jsel || !idx || !xyz || !new_xyz || !coef || !dPf || !gstats) return PCM_EINVAL;
    if (H % 4 || H > 4096) return PCM_EUNSUPPORTED;
    sa_bwd_scatter_kernel<<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(dout, out, jsel, idx, xyz, new_xyz, coef, m, k, H,
                                                                                  dPf, gstats, 0, 0, 0, nullptr);
    return pcm_launch_status();
}

PCM_API int pcm_sa_output_tokens(int m, int H, int per_cloud, int batch, int head_rows, const float* ymax, const float* ymin,
                                 const unsigned char* jmax, const unsigned char* jmin, const float* coef, const float* pos,
                                 float* out, void* out_bf16, void* out_pos_bf16, unsigned char* jsel, pcm_stream_t stream) {
    const long total = (long)m * H;
    if (total <= 0) return PCM_OK;
    if (!ymax || !jmax || !coef || !out || !jsel || (!ymin != !jmin) || (out_pos_bf16 && !pos)) return PCM_EINVAL;
    if (H % 4 || per_cloud <= 0 || batch <= 0 || head_rows < 0 || (long)per_cloud * batch != m) return PCM_EUNSUPPORTED;
    sa_output_tokens_kernel<<<sa_grid((total / 4 + 255) / 256), 256, 0, pcm_cu_stream(stream)>>>(
        ymax, ymin, jmax, jmin, coef, total, H, per_cloud, batch, head_rows, pos, out, reinterpret_cast<__nv_bfloat16*>(out_bf16),
        reinterpret_cast<__nv_bfloat16*>(out_pos_bf16), jsel);
    return pcm_launch_status();
}

PCM_API int pcm_sa_bwd_scatter_tokens(int m, int k, int H, int per_cloud, int batch, int head_rows, const float* dout,
                                      const float* dout2, const float* out, const unsigned char* jsel, const int* idx,
                                      const float* xyz, const float* new_xyz, const float* coef, float* dPf, double* gstats,
                                      pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (!dout || !out || !jsel || !idx || !xyz || !new_xyz || !coef || !dPf || !gstats) return PCM_EINVAL;
    if (H % 4 || H > 4096 || per_cloud <= 0 || batch <= 0 || head_rows < 0 || (long)per_cloud * batch != m) return PCM_EUNSUPPORTED;
    sa_bwd_scatter_kernel<<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(dout, out, jsel, idx, xyz, new_xyz, coef, m, k, H,
                                                                                  dPf, gstats, per_cloud, batch, head_rows, dout2);
    return pcm_launch_status();
}

// Cloud-slice form of pcm_sa_gather_stats (see sa_gather_stats_cloud_kernel): the caller also passes the cumulative cloud
// offsets of the source points and of the queries (b clouds) and a host-known upper bound n_max of the cloud sizes.
// Returns PCM_EUNSUPPORTED when the shape is outside the fast path (k != 16, H % 4, cloud too large for shared memory):
// call the generic entry point then.  (The same layout was tried for the backward scatter -- sparse gradient accumulated
// with shared-memory atomics, dPf written once -- and measured SLOWER than the global-atomic kernel: 245 vs 194 us at cfg-2.)
static int sa_gather_clouds_impl(int b, int n_max, int m, int k, int H, const float* Pf, const float* xyz, const float* new_xyz,
                                 const int* idx, const int* offset, const int* new_offset, const float* W, int ldw,
                                 const float* sel_gamma, float* ymax, float* ymin, unsigned char* jmax, unsigned char* jmin,
                                 double* stats, pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (!Pf || !xyz || !new_xyz || !idx || !offset || !new_offset || !W || !ymax || !jmax || !stats) return PCM_EINVAL;
    if (!sel_gamma && (!ymin || !jmin)) return PCM_EINVAL;
    if (b <= 0 || n_max <= 0) return PCM_EINVAL;
    const int cs = k == 16 ? sa_slice_width(n_max, H) : 0;
    if (!cs) return PCM_EUNSUPPORTED;
    const size_t smem = (size_t)n_max * cs * 4;
    const dim3 grid(H / cs, b);
    cudaStream_t st = pcm_cu_stream(stream);
    cudaError_t e = cudaSuccess;
#define PCM_SA_FWD(CSV, ONE)                                                                                                  \
    e = sa_allow_smem(sa_gather_stats_cloud_kernel<CSV, 16, ONE>, smem);                                                      \
    if (e == cudaSuccess)                                                                                                     \
        sa_gather_stats_cloud_kernel<CSV, 16, ONE><<<grid, 256, smem, st>>>(Pf, xyz, new_xyz, idx, offset, new_offset, W, ldw, \
                                                                           sel_gamma, H, n_max, ymax, ymin, jmax, jmin, stats)
    if (sel_gamma) {
        switch (cs) {
            case 32: PCM_SA_FWD(32, true); break;
            case 16: PCM_SA_FWD(16, true); break;
            case 8: PCM_SA_FWD(8, true); break;
            default: PCM_SA_FWD(4, true); break;
        }
    } else {
        switch (cs) {
            case 32: PCM_SA_FWD(32, false); break;
            case 16: PCM_SA_FWD(16, false); break;
            case 8: PCM_SA_FWD(8, false); break;
            default: PCM_SA_FWD(4, false); break;
        }
    }
#undef PCM_SA_FWD
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_sa_gather_stats_clouds(int b, int n_max, int m, int k, int H, const float* Pf, const float* xyz,
                                       const float* new_xyz, const int* idx, const int* offset, const int* new_offset,
                                       const float* W, int ldw, float* ymax, float* ymin, unsigned char* jmax,
                                       unsigned char* jmin, double* stats, pcm_stream_t stream) {
    return sa_gather_clouds_impl(b, n_max, m, k, H, Pf, xyz, new_xyz, idx, offset, new_offset, W, ldw, nullptr, ymax, ymin, jmax,
                                 jmin, stats, stream);
}

// Single-extreme forms of the gather pass.  BatchNorm's per-channel scale a_c = gamma_c * invstd_c has the sign of gamma_c,
// which is known BEFORE the batch statistics: only the extreme that will be selected is tracked (max of y where gamma >= 0,
// min where gamma < 0) -- a third less compare / select work per edge and half the (m, H) intermediates.  yext / jext take
// the place of ymax / jmax in pcm_sa_output[_tokens] (pass ymin = jmin = NULL there).  Bit-identical results.
// pcm_sa_gather_sel: nsample 16 or 32 (PCM_EUNSUPPORTED otherwise); pcm_sa_gather_sel_clouds: as pcm_sa_gather_stats_clouds.
PCM_API int pcm_sa_gather_sel(int m, int k, int H, const float* Pf, const float* xyz, const float* new_xyz, const int* idx,
                              const float* W, int ldw, const float* gamma, float* yext, unsigned char* jext, double* stats,
                              pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (!Pf || !xyz || !new_xyz || !idx || !W || !gamma || !yext || !jext || !stats) return PCM_EINVAL;
    if (H % 4 || H > 4096) return PCM_EUNSUPPORTED;
    if (k == 16)
        sa_gather_stats_k_kernel<16, true><<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(Pf, xyz, new_xyz, idx, W, ldw, gamma, m, H,
                                                                                                   yext, nullptr, jext, nullptr, stats);
    else if (k == 32)
        sa_gather_stats_k_kernel<32, true><<<sa_grid(m), sa_threads(H), 0, pcm_cu_stream(stream)>>>(Pf, xyz, new_xyz, idx, W, ldw, gamma, m, H,
                                                                                                   yext, nullptr, jext, nullptr, stats);
    else
        return PCM_EUNSUPPORTED;
    return pcm_launch_status();
}

PCM_API int pcm_sa_gather_sel_clouds(int b, int n_max, int m, int k, int H, const float* Pf, const float* xyz,
                                     const float* new_xyz, const int* idx, const int* offset, const int* new_offset,
                                     const float* W, int ldw, const float* gamma, float* yext, unsigned char* jext,
                                     double* stats, pcm_stream_t stream) {
    if (!gamma) return PCM_EINVAL;
    return sa_gather_clouds_impl(b, n_max, m, k, H, Pf, xyz, new_xyz, idx, offset, new_offset, W, ldw, gamma, yext, nullptr, jext,
                                 nullptr, stats, stream);
}

PCM_API int pcm_sa_edge_stats(int m, int k, const int* idx, const float* xyz, const float* new_xyz, float* cnt,
                              float* sq, double* sdtot, pcm_stream_t stream) {
    const long edges = (long)m * k;
    if (edges <= 0) return PCM_OK;
    if (!idx || !xyz || !new_xyz || !cnt || !sq || !sdtot) return PCM_EINVAL;
    sa_edge_stats_kernel<<<sa_grid((edges + 255) / 256), 256, 0, pcm_cu_stream(stream)>>>(idx, xyz, new_xyz, edges, k, cnt, sq, sdtot);
    return pcm_launch_status();
}

PCM_API int pcm_sa_bwd_coef(int H, const double* gstats, const double* fstats, const double* sdtot, const float* coef,
                            double n_rows, int training, float* ab, float* dW, int ldw, float* dgamma, float* dbeta,
                            pcm_stream_t stream) {
    return pcm_sa_bwd_coef_ex(H, gstats, fstats, sdtot, coef, n_rows, nullptr, training, ab, dW, ldw, dgamma, dbeta, stream);
}

PCM_API int pcm_sa_bwd_coef_ex(int H, const double* gstats, const double* fstats, const double* sdtot, const float* coef,
                               double n_rows, const double* n_rows_dev, int training, float* ab, float* dW, int ldw,
                               float* dgamma, float* dbeta, pcm_stream_t stream) {
    if (H <= 0) return PCM_OK;
    if (!gstats || !fstats || !sdtot || !coef || !ab) return PCM_EINVAL;
    sa_bwd_coef_kernel<<<pcm_divup(H, 128), 128, 0, pcm_cu_stream(stream)>>>(gstats, fstats, sdtot, coef, n_rows, training, H, ab, dW,
                                                                            ldw, dgamma, dbeta, n_rows_dev);
    return pcm_launch_status();
}

PCM_API int pcm_sa_bwd_dense(int n, int H, const float* Pf, const float* xyz, const float* cnt, const float* sq,
                             const float* W, int ldw, const float* ab, const float* dPf, void* dPf_bf16,
                             pcm_stream_t stream) {
    if (n <= 0) return PCM_OK;
    if (!Pf || !xyz || !cnt || !sq || !W || !ab || !dPf || !dPf_bf16) return PCM_EINVAL;
    if (H % 4 || H > 4096) return PCM_EUNSUPPORTED;
    sa_bwd_dense_kernel<<<sa_grid(n), sa_threads(H), 0, pcm_cu_stream(stream)>>>(Pf, xyz, cnt, sq, W, ldw, ab, n, H, dPf,
                                                                                reinterpret_cast<__nv_bfloat16*>(dPf_bf16));
    return pcm_launch_status();
}
