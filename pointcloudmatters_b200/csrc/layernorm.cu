// layernorm.cu -- fused residual + dropout + LayerNorm (forward / backward) and column sums, sm_100a.
//
// Every transformer sub-block of the reference ends with `norm(x_res + dropout(sublayer))`
// (transformer.py:249-253,333-345): three ATen kernels forward and four backward (whose
// gamma/beta reduction alone was 10% of the step in the ncu launch list).  Here:
//   forward : h = res + dropout(x);  y = (h - mean) * rstd * gamma + beta   -> y (fp32), h, mean, rstd
//   backward: dh = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;
//             dx = dropout-backward(dh); dgamma += dy * xhat; dbeta += dy
// One warp owns one row (C = 128 * V floats, V float4 per lane, fully coalesced 128-bit accesses),
// row statistics by warp shuffles, dgamma / dbeta accumulated in registers across the rows a warp
// visits and flushed with one fp32 atomic per column per CTA.  HBM-bound: 12-16 B per element.
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PCM_FULL_MASK, v, o);
    return v;
}

// Register budget: the kernel is HBM-bound and its only latency hiding is warps in flight (each warp has 2 V 16-byte loads
// outstanding, then two dependent shuffle reductions, then the stores).  gamma / beta are therefore NOT kept in registers
// (2 V float4 per thread) but re-read per row -- L1 hits -- which brings V = 4 from 126 to <= 64 registers: 4 CTAs (32 warps)
// per SM instead of 2 (ncu r1: 59 % of the HBM peak at 25 % warps active).
template <int V>
__global__ void __launch_bounds__(256, V <= 4 ? 4 : 1) add_dropout_ln_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, long rows, float eps, float p_drop, const unsigned long long* __restrict__ seed_base,
    unsigned long long seed_offset, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ h_out,
    float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ pos, int pos_row_div,
    __nv_bfloat16* __restrict__ ypos_bf16) {
    constexpr int C = 128 * V;
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const uint32_t thr16 = pcm_drop_thr16(p_drop);
    const float ks = p_drop > 0.f ? pcm_keep_scale(thr16) : 1.0f;
    for (long r = wid; r < rows; r += nwarps) {
        float4 h[V];
        float s = 0.f;
        const uint32_t rseed = pcm_row_seed(seed, (unsigned long long)r * C);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const size_t e = (size_t)r * C + (size_t)(v * 32 + lane) * 4;
            float4 xv = x ? *reinterpret_cast<const float4*>(x + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (p_drop > 0.f) {
                const uint32_t c = (uint32_t)(v * 32 + lane);
                const uint32_t h0 = pcm_pair_bits(rseed, 2 * c), h1 = pcm_pair_bits(rseed, 2 * c + 1);
                xv.x = (h0 & 0xFFFFu) >= thr16 ? xv.x * ks : 0.f;
                xv.y = (h0 >> 16) >= thr16 ? xv.y * ks : 0.f;
                xv.z = (h1 & 0xFFFFu) >= thr16 ? xv.z * ks : 0.f;
                xv.w = (h1 >> 16) >= thr16 ? xv.w * ks : 0.f;
            }
            const float4 rv = *reinterpret_cast<const float4*>(res + e);
            h[v] = make_float4(rv.x + xv.x, rv.y + xv.y, rv.z + xv.z, rv.w + xv.w);
            s += (h[v].x + h[v].y) + (h[v].z + h[v].w);
        }
        const float mean = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float a = h[v].x - mean, b = h[v].y - mean, c = h[v].z - mean, d = h[v].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const size_t e = (size_t)r * C + (size_t)(v * 32 + lane) * 4;
            const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
            const float4 bv = __ldg(reinterpret_cast<const float4*>(beta) + v * 32 + lane);
            float4 o;
            o.x = (h[v].x - mean) * rstd * gv.x + bv.x;
            o.y = (h[v].y - mean) * rstd * gv.y + bv.y;
            o.z = (h[v].z - mean) * rstd * gv.z + bv.z;
            o.w = (h[v].w - mean) * rstd * gv.w + bv.w;
            *reinterpret_cast<float4*>(y + e) = o;
            if (y_bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(y_bf16 + e) = pk;
            }
            if (ypos_bf16) {  // operand of the next attention block: bf16(y + pos), pos row-broadcast over the batch
                const float4 pv = *reinterpret_cast<const float4*>(pos + (size_t)(r / pos_row_div) * C + (size_t)(v * 32 + lane) * 4);
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x + pv.x, o.y + pv.y), hi = __floats2bfloat162_rn(o.z + pv.z, o.w + pv.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(ypos_bf16 + e) = pk;
            }
            if (h_out) *reinterpret_cast<float4*>(h_out + e) = h[v];
        }
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
    }
}

template <int V, bool CS>
__global__ void __launch_bounds__(256, V <= 4 ? 2 : 1) add_dropout_ln_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ dy_b, const float* __restrict__ h, const float* __restrict__ mean_in,
    const float* __restrict__ rstd_in, const float* __restrict__ gamma, long rows, float p_drop,
    const unsigned long long* __restrict__ seed_base, unsigned long long seed_offset, float* __restrict__ dres,
    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx_bf16,
    float* __restrict__ dx_colsum) {
    // dx_colsum (optional, C floats, accumulated): column sums of dx = the bias gradient of the linear layer that produced
    // x (out_proj / linear2) -- formed here, where dx is in registers, instead of by a colsum pass re-reading dx
    constexpr int C = 128 * V;
    __shared__ float sg[C], sb[C];
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const uint32_t thr16 = pcm_drop_thr16(p_drop);
    const float ks = p_drop > 0.f ? pcm_keep_scale(thr16) : 1.0f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { sg[i] = 0.f; sb[i] = 0.f; }
    __syncthreads();
    // register budget: 2 CTAs per SM need <= 128 registers per thread; with the third accumulator set (CS) gamma is
    // re-read per row (an L1 hit) instead of living in registers
    float4 g4[CS ? 1 : V], ag[V], abt[V], axs[CS ? V : 1];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (!CS) g4[v] = reinterpret_cast<const float4*>(gamma)[v * 32 + lane];
        ag[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        abt[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CS) axs[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long r = wid; r < rows; r += nwarps) {
        const float mean = mean_in[r], rstd = rstd_in[r];
        const uint32_t rseed = pcm_row_seed(seed, (unsigned long long)r * C);
        float4 gy[V], xh[V];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const size_t e = (size_t)r * C + (size_t)(v * 32 + lane) * 4;
            float4 d = *reinterpret_cast<const float4*>(dy + e);
            if (dy_b) {  // second gradient contribution (the residual branch of the consumer): summed on load
                const float4 d2 = *reinterpret_cast<const float4*>(dy_b + e);
                d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
            }
            const float4 hv = *reinterpret_cast<const float4*>(h + e);
            xh[v] = make_float4((hv.x - mean) * rstd, (hv.y - mean) * rstd, (hv.z - mean) * rstd, (hv.w - mean) * rstd);
            const float4 gm = CS ? __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane) : g4[CS ? 0 : v];
            gy[v] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
            s1 += (gy[v].x + gy[v].y) + (gy[v].z + gy[v].w);
            s2 += (gy[v].x * xh[v].x + gy[v].y * xh[v].y) + (gy[v].z * xh[v].z + gy[v].w * xh[v].w);
            ag[v].x += d.x * xh[v].x; ag[v].y += d.y * xh[v].y; ag[v].z += d.z * xh[v].z; ag[v].w += d.w * xh[v].w;
            abt[v].x += d.x; abt[v].y += d.y; abt[v].z += d.z; abt[v].w += d.w;
        }
        const float c1 = warp_sum(s1) * (1.0f / C), c2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const size_t e = (size_t)r * C + (size_t)(v * 32 + lane) * 4;
            float4 dh;
            dh.x = rstd * (gy[v].x - c1 - xh[v].x * c2);
            dh.y = rstd * (gy[v].y - c1 - xh[v].y * c2);
            dh.z = rstd * (gy[v].z - c1 - xh[v].z * c2);
            dh.w = rstd * (gy[v].w - c1 - xh[v].w * c2);
            if (dres) *reinterpret_cast<float4*>(dres + e) = dh;
            float4 o = dh;
            if ((dx || dx_bf16) && dx != dres) {  // dx == NULL with dx_bf16: only the bf16 copy of dx is wanted
                if (p_drop > 0.f) {
                    const uint32_t c = (uint32_t)(v * 32 + lane);
                    const uint32_t h0 = pcm_pair_bits(rseed, 2 * c), h1 = pcm_pair_bits(rseed, 2 * c + 1);
                    o.x = (h0 & 0xFFFFu) >= thr16 ? dh.x * ks : 0.f;
                    o.y = (h0 >> 16) >= thr16 ? dh.y * ks : 0.f;
                    o.z = (h1 & 0xFFFFu) >= thr16 ? dh.z * ks : 0.f;
                    o.w = (h1 >> 16) >= thr16 ? dh.w * ks : 0.f;
                }
                if (dx) *reinterpret_cast<float4*>(dx + e) = o;
            }
            if (CS) { axs[v].x += o.x; axs[v].y += o.y; axs[v].z += o.z; axs[v].w += o.w; }
            if (dx_bf16) {  // bf16 copy of dx: the operand of the sub-block's backward GEMMs
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dx_bf16 + e) = pk;
            }
        }
    }
    // flush per-warp partials: shared-memory atomics per CTA, then one global atomic per column
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int c = (v * 32 + lane) * 4;
        atomicAdd(&sg[c + 0], ag[v].x); atomicAdd(&sg[c + 1], ag[v].y); atomicAdd(&sg[c + 2], ag[v].z); atomicAdd(&sg[c + 3], ag[v].w);
        atomicAdd(&sb[c + 0], abt[v].x); atomicAdd(&sb[c + 1], abt[v].y); atomicAdd(&sb[c + 2], abt[v].z); atomicAdd(&sb[c + 3], abt[v].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(dgamma + i, sg[i]);
        atomicAdd(dbeta + i, sb[i]);
    }
    if (CS) {  // third per-column reduction reuses sg after the flush above
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) sg[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int c = (v * 32 + lane) * 4;
            atomicAdd(&sg[c + 0], axs[v].x); atomicAdd(&sg[c + 1], axs[v].y); atomicAdd(&sg[c + 2], axs[v].z); atomicAdd(&sg[c + 3], axs[v].w);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dx_colsum + i, sg[i]);
    }
}

// out[c] += sum_r src[r, c]; src fp32 or bf16 with row pitch ld.
// Vector form (C % VEC == 0, 16-byte aligned rows): a thread owns VEC consecutive columns and
// walks rows with 16-byte loads, several row groups per CTA in flight; partial sums meet in
// shared memory and leave as one atomic per column per CTA.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ src, long rows, int C, long ld,
                                                         int rows_per_cta, float* __restrict__ out) {
    extern __shared__ float part[];  // [groups][C]
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const int tpr = C / VEC;                      // threads per row
    const int groups = max(1, 256 / tpr);          // row groups per CTA
    const int rg = threadIdx.x / tpr, tc = threadIdx.x - rg * tpr;
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (rg < groups) {
        const int c0 = tc * VEC;  // tpr * VEC == C: one column chunk per thread
#pragma unroll 4
        for (long r = r0 + rg; r < r1; r += groups) {
            const uint4 raw = *reinterpret_cast<const uint4*>(src + r * ld + c0);
            if (sizeof(T) == 2) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
            } else {
                const float* f = reinterpret_cast<const float*>(&raw);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] += f[k];
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) part[rg * C + c0 + k] = acc[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int g = 0; g < groups; ++g) t += part[g * C + c];
        atomicAdd(out + c, t);
    }
}

// Grouped form: up to CS_MAX independent bf16 column-sum problems (bias gradients of the in-projections: colsum of the
// [dQ | dK | dV] buffers) in ONE launch.  They are off the backward's critical path, so the operator layer queues them
// next to the weight-gradient GEMMs (functional.DW_QUEUE) instead of launching ~30 latency-bound kernels per step.
constexpr int CS_MAX = 64;
struct ColsumProblem { const __nv_bfloat16* src; float* out; long rows, ld; int C, rows_per_cta; };
struct ColsumGroup { int n; int cta_prefix[CS_MAX + 1]; ColsumProblem prob[CS_MAX]; };

__global__ void __launch_bounds__(256) colsum_grouped_kernel(const __grid_constant__ ColsumGroup g) {
    extern __shared__ float part[];
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    int lo = 0, hi = g.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (g.cta_prefix[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const ColsumProblem& q = g.prob[lo];
    const int C = q.C, tpr = C / 8, groups = max(1, 256 / tpr);
    const int rg = threadIdx.x / tpr, tc = threadIdx.x - rg * tpr;
    const long r0 = (long)((int)blockIdx.x - g.cta_prefix[lo]) * q.rows_per_cta;
    const long r1 = r0 + q.rows_per_cta < q.rows ? r0 + q.rows_per_cta : q.rows;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rg < groups) {
        const int c0 = tc * 8;
#pragma unroll 4
        for (long r = r0 + rg; r < r1; r += groups) {
            const uint4 raw = *reinterpret_cast<const uint4*>(q.src + r * q.ld + c0);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) part[rg * C + c0 + k] = acc[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int gg = 0; gg < groups; ++gg) t += part[gg * C + c];
        atomicAdd(q.out + c, t);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ src, long rows, int C, long ld, int rows_per_cta,
                                                     float* __restrict__ out) {
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (long r = r0; r < r1; ++r) acc += (float)src[r * ld + c];
        atomicAdd(out + c, acc);
    }
}

// out_bf16[r, c] = bf16(a[r, c] + b[r / b_row_div, c])  (b may be NULL; b_row_div > 1 broadcasts a
// (L, 1, C) positional table over the batch of token-major (L*B, C) activations)
__global__ void __launch_bounds__(256) add_cast_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            long n4, int C4, int b_row_div, __nv_bfloat16* __restrict__ out) {
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(a)[i];
        if (b) {
            long bi = i;
            if (b_row_div > 1) { const long r = i / C4; bi = (r / b_row_div) * C4 + (i - r * C4); }
            const float4 w = reinterpret_cast<const float4*>(b)[bi];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
}

// ---- FFN hidden layer (dim_feedforward = 32 in the reference configs): dropout of the ReLU'd hidden activation
// and its backward.  The hidden tensor is (rows, Hd) bf16 straight out of the first GEMM's bias+ReLU epilogue;
// one thread handles 8 consecutive elements (16-byte accesses), masks come from the counter-based RNG shared with
// the LayerNorm kernels (stateless: the backward regenerates them).  Replaces six ATen kernels per FFN
// (fused_dropout, bf16 cast, masked_scale, compare, multiply, bf16 cast) with two.
__device__ __forceinline__ void ffn_keep8(uint32_t rseed, uint32_t chunk, uint32_t thr16, bool keep[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t hbits = pcm_pair_bits(rseed, chunk * 4 + k);
        keep[2 * k] = (hbits & 0xFFFFu) >= thr16;
        keep[2 * k + 1] = (hbits >> 16) >= thr16;
    }
}

__global__ void __launch_bounds__(256)
ffn_dropout_fwd_kernel(const __nv_bfloat16* __restrict__ h, long rows, int Hd, float p_drop,
                       const unsigned long long* __restrict__ seed_base, unsigned long long seed_offset,
                       __nv_bfloat16* __restrict__ out) {
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const uint32_t thr16 = pcm_drop_thr16(p_drop);
    const float ks = pcm_keep_scale(thr16);
    const int cpr = Hd / 8;  // 16-byte chunks per row
    const long total = rows * cpr;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / cpr;
        const uint32_t c = (uint32_t)(i - r * cpr);
        const uint4 raw = reinterpret_cast<const uint4*>(h)[i];
        const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
        bool keep[8];
        ffn_keep8(pcm_row_seed(seed, (unsigned long long)r * Hd), c, thr16, keep);
        uint4 o;
        uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(v[k]);
            __nv_bfloat162 w = __floats2bfloat162_rn(keep[2 * k] ? f.x * ks : 0.f, keep[2 * k + 1] ? f.y * ks : 0.f);
            op[k] = *reinterpret_cast<uint32_t*>(&w);
        }
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// dh = bf16( d(dropped hidden) * keep * scale * [h > 0] ); p_drop = 0 -> ReLU gate only.
__global__ void __launch_bounds__(256)
ffn_relu_dropout_bwd_kernel(const float* __restrict__ dhd, const __nv_bfloat16* __restrict__ h, long rows, int Hd,
                            float p_drop, const unsigned long long* __restrict__ seed_base,
                            unsigned long long seed_offset, __nv_bfloat16* __restrict__ dh, float* __restrict__ dh_colsum) {
    // dh_colsum (optional, Hd floats, accumulated) = column sums of dh = the gradient of linear1's bias; the launcher passes
    // it only when every thread keeps the same 8-column chunk over its grid-stride loop ((grid * block) % (Hd / 8) == 0)
    __shared__ float scs[256];
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const uint32_t thr16 = pcm_drop_thr16(p_drop);
    const float ks = p_drop > 0.f ? pcm_keep_scale(thr16) : 1.0f;
    const int cpr = Hd / 8;
    const long total = rows * cpr;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / cpr;
        const uint32_t c = (uint32_t)(i - r * cpr);
        const uint4 raw = reinterpret_cast<const uint4*>(h)[i];
        const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float4 g0 = reinterpret_cast<const float4*>(dhd)[2 * i], g1 = reinterpret_cast<const float4*>(dhd)[2 * i + 1];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        bool keep[8];
        if (p_drop > 0.f) {
            ffn_keep8(pcm_row_seed(seed, (unsigned long long)r * Hd), c, thr16, keep);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) keep[k] = true;
        }
        uint4 o;
        uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(v[k]);
            const float w0 = (keep[2 * k] && f.x > 0.f) ? g[2 * k] * ks : 0.f;
            const float w1 = (keep[2 * k + 1] && f.y > 0.f) ? g[2 * k + 1] * ks : 0.f;
            cs[2 * k] += w0;
            cs[2 * k + 1] += w1;
            __nv_bfloat162 w = __floats2bfloat162_rn(w0, w1);
            op[k] = *reinterpret_cast<uint32_t*>(&w);
        }
        reinterpret_cast<uint4*>(dh)[i] = o;
    }
    if (dh_colsum) {
        for (int i = threadIdx.x; i < Hd; i += blockDim.x) scs[i] = 0.f;
        __syncthreads();
        const int c = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) % cpr) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&scs[c + k], cs[k]);
        __syncthreads();
        for (int i = threadIdx.x; i < Hd; i += blockDim.x) atomicAdd(dh_colsum + i, scs[i]);
    }
}

int g_ln_fwd_cap = 148 * 8;

inline int ln_grid(long rows) {
    long blocks = (rows + 7) / 8;
    // one resident wave (4 CTAs per SM): tools/bench_ln.py, 6 400 rows: 21.1 / 14.3 / 12.7 / 14.5 us at 148 / 296 / 592 / 1 184
    // CTAs, 32 960 rows: 119 / 79 / 61.3 / 61.5 us
    const long cap = g_ln_fwd_cap > 0 ? g_ln_fwd_cap : 148L * 4;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// Launch-shape knobs of the two kernels that end in one global atomic per column per CTA (LayerNorm backward:
// dgamma / dbeta, 2C atomics per CTA; colsum: C per CTA).  With one row per warp a 6 400-row backward ran 800 CTAs
// = 819 200 atomics onto 1 024 addresses and was atomic-bound (21 us against a 9 us HBM floor); fewer, longer-lived
// CTAs trade a little memory-level parallelism for a 3-8x shorter atomic tail.  Measured with tools/bench_ln.py
// (CUDA-graph replay, one B200, profiles/r1_ln_colsum_sweep.jsonl): backward, 6 400 rows: 20.6 us at 800 CTAs ->
// 12.0 us at 296 (2 per SM); 32 960 rows: 57.5 us at 1 184 -> 52.2 us at 222; colsum, 6 400 rows: 5.5 us at 16
// rows per CTA -> 3.9 us at >= 32, 32 960 rows best with ~592 CTAs (11.8 us).
int g_ln_bwd_cap = 0;  // 0 = 2 CTAs per SM (222 vs 296 at 32 960 rows is within run-to-run noise: 52-57 us)
int g_colsum_ctas = 148 * 4;
int g_colsum_min_rows = 32;

inline int ln_grid_bwd(long rows) {
    long blocks = (rows + 7) / 8;
    const long cap = g_ln_bwd_cap > 0 ? g_ln_bwd_cap : 296L;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

#define LN_DISPATCH(V, KERNEL, ...)                                   \
    switch (V) {                                                      \
        case 1: pcm_launch(KERNEL<1>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        case 2: pcm_launch(KERNEL<2>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        case 4: pcm_launch(KERNEL<4>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        case 8: pcm_launch(KERNEL<8>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        default: return PCM_EUNSUPPORTED;                             \
    }

#define LN_DISPATCH2(V, KERNEL, FLAG, ...)                            \
    switch (V) {                                                      \
        case 1: pcm_launch(KERNEL<1, FLAG>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        case 2: pcm_launch(KERNEL<2, FLAG>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        case 4: pcm_launch(KERNEL<4, FLAG>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        case 8: pcm_launch(KERNEL<8, FLAG>, dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;  \
        default: return PCM_EUNSUPPORTED;                             \
    }

// y = LayerNorm(res + dropout(x)) (x may be NULL = plain LayerNorm(res)); C in {128, 256, 512, 1024}.
// Optional outputs: y_bf16 (operand of the next GEMM), h (= res + dropout(x), saved for backward),
// mean / rstd (rows).
PCM_API int pcm_add_dropout_ln_fwd_ex(long long rows, int C, const float* x, const float* res, const float* gamma,
                                      const float* beta, float eps, float p_drop, const unsigned long long* seed_base,
                                      unsigned long long seed_offset, float* y, void* y_bf16, float* h, float* mean,
                                      float* rstd, const float* pos, int pos_row_div, void* ypos_bf16,
                                      pcm_stream_t stream) {
    if (rows <= 0) return PCM_OK;
    if (!res || !gamma || !beta || !y || (C % 128)) return (C % 128) ? PCM_EUNSUPPORTED : PCM_EINVAL;
    if (ypos_bf16 && (!pos || pos_row_div < 1)) return PCM_EINVAL;
    cudaStream_t st = pcm_cu_stream(stream);
    const int grid = ln_grid(rows);
    LN_DISPATCH(C / 128, add_dropout_ln_fwd_kernel, x, res, gamma, beta, rows, eps, p_drop, seed_base, seed_offset, y,
                reinterpret_cast<__nv_bfloat16*>(y_bf16), h, mean, rstd, pos, pos_row_div,
                reinterpret_cast<__nv_bfloat16*>(ypos_bf16))
    return pcm_launch_status();
}

PCM_API int pcm_add_dropout_ln_fwd(long long rows, int C, const float* x, const float* res, const float* gamma,
                                   const float* beta, float eps, float p_drop, const unsigned long long* seed_base,
                                   unsigned long long seed_offset, float* y, void* y_bf16, float* h, float* mean,
                                   float* rstd, pcm_stream_t stream) {
    return pcm_add_dropout_ln_fwd_ex(rows, C, x, res, gamma, beta, eps, p_drop, seed_base, seed_offset, y, y_bf16, h, mean,
                                     rstd, nullptr, 1, nullptr, stream);
}

// dres = dLN/dh; dx = dropout-backward(dres) (pass dx == dres or NULL when not needed);
// dgamma / dbeta are ACCUMULATED (caller zero-fills); dx_bf16 (optional) = bf16(dx).  dx == NULL with dx_bf16 given:
// only the bf16 copy is written (the consumer is a GEMM; saves the fp32 store, 1/6 of the kernel's traffic).
PCM_API int pcm_add_dropout_ln_bwd_ex2(long long rows, int C, const float* dy, const float* dy_b, const float* h, const float* mean,
                                       const float* rstd, const float* gamma, float p_drop,
                                       const unsigned long long* seed_base, unsigned long long seed_offset, float* dres,
                                       float* dx, float* dgamma, float* dbeta, void* dx_bf16, float* dx_colsum,
                                       pcm_stream_t stream) {
    if (rows <= 0) return PCM_OK;
    if (!dy || !h || !mean || !rstd || !gamma || !dgamma || !dbeta) return PCM_EINVAL;
    if (dx_colsum && !dx && !dx_bf16) return PCM_EINVAL;
    if (C % 128) return PCM_EUNSUPPORTED;
    cudaStream_t st = pcm_cu_stream(stream);
    const int grid = ln_grid_bwd(rows);
    if (dx_colsum) {
        LN_DISPATCH2(C / 128, add_dropout_ln_bwd_kernel, true, dy, dy_b, h, mean, rstd, gamma, rows, p_drop, seed_base, seed_offset,
                     dres, dx, dgamma, dbeta, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dx_colsum)
    } else {
        LN_DISPATCH2(C / 128, add_dropout_ln_bwd_kernel, false, dy, dy_b, h, mean, rstd, gamma, rows, p_drop, seed_base, seed_offset,
                     dres, dx, dgamma, dbeta, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dx_colsum)
    }
    return pcm_launch_status();
}

PCM_API int pcm_add_dropout_ln_bwd_ex(long long rows, int C, const float* dy, const float* dy_b, const float* h, const float* mean,
                                      const float* rstd, const float* gamma, float p_drop,
                                      const unsigned long long* seed_base, unsigned long long seed_offset, float* dres,
                                      float* dx, float* dgamma, float* dbeta, void* dx_bf16, pcm_stream_t stream) {
    return pcm_add_dropout_ln_bwd_ex2(rows, C, dy, dy_b, h, mean, rstd, gamma, p_drop, seed_base, seed_offset, dres, dx, dgamma,
                                      dbeta, dx_bf16, nullptr, stream);
}

PCM_API int pcm_add_dropout_ln_bwd(long long rows, int C, const float* dy, const float* h, const float* mean,
                                   const float* rstd, const float* gamma, float p_drop,
                                   const unsigned long long* seed_base, unsigned long long seed_offset, float* dres,
                                   float* dx, float* dgamma, float* dbeta, pcm_stream_t stream) {
    return pcm_add_dropout_ln_bwd_ex(rows, C, dy, nullptr, h, mean, rstd, gamma, p_drop, seed_base, seed_offset, dres, dx, dgamma,
                                     dbeta, nullptr, stream);
}

PCM_API int pcm_ffn_dropout_fwd(long long rows, int Hd, const void* h, float p_drop, const unsigned long long* seed_base,
                                unsigned long long seed_offset, void* out, pcm_stream_t stream) {
    if (rows <= 0 || Hd <= 0) return PCM_OK;
    if (!h || !out || p_drop <= 0.f || p_drop >= 1.f) return PCM_EINVAL;
    if (Hd % 8) return PCM_EUNSUPPORTED;
    const long total = (long)rows * (Hd / 8);
    const int grid = (int)((total + 255) / 256 < 148L * 8 ? (total + 255) / 256 : 148L * 8);
    cudaError_t e = pcm_launch(ffn_dropout_fwd_kernel, dim3(grid), dim3(256), 0, pcm_cu_stream(stream),
                               reinterpret_cast<const __nv_bfloat16*>(h), (long)rows, Hd, p_drop, seed_base, seed_offset,
                               reinterpret_cast<__nv_bfloat16*>(out));
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_ffn_relu_dropout_bwd(long long rows, int Hd, const float* dhd, const void* h, float p_drop,
                                     const unsigned long long* seed_base, unsigned long long seed_offset, void* dh,
                                     pcm_stream_t stream) {
    return pcm_ffn_relu_dropout_bwd_ex(rows, Hd, dhd, h, p_drop, seed_base, seed_offset, dh, nullptr, stream);
}

PCM_API int pcm_ffn_relu_dropout_bwd_ex(long long rows, int Hd, const float* dhd, const void* h, float p_drop,
                                        const unsigned long long* seed_base, unsigned long long seed_offset, void* dh,
                                        float* dh_colsum, pcm_stream_t stream) {
    if (rows <= 0 || Hd <= 0) return PCM_OK;
    if (!dhd || !h || !dh || p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    if (Hd % 8) return PCM_EUNSUPPORTED;
    if (dh_colsum && (Hd > 256 || 256 % (Hd / 8))) return PCM_EUNSUPPORTED;  // see the kernel: fixed chunk per thread
    const long total = (long)rows * (Hd / 8);
    // with the column sums every CTA ends in Hd global atomics onto the same Hd addresses: one CTA per SM keeps that tail short
    const long cap = dh_colsum ? 148L : 148L * 8;
    const int grid = (int)((total + 255) / 256 < cap ? (total + 255) / 256 : cap);
    cudaError_t e = pcm_launch(ffn_relu_dropout_bwd_kernel, dim3(grid), dim3(256), 0, pcm_cu_stream(stream), dhd,
                               reinterpret_cast<const __nv_bfloat16*>(h), (long)rows, Hd, p_drop, seed_base, seed_offset,
                               reinterpret_cast<__nv_bfloat16*>(dh), dh_colsum);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

// Debug aid for tools/bench_ln.py: override the launch-shape knobs above (values <= 0 keep the current one).
PCM_API int pcm_ln_debug_tune(int ln_fwd_max_ctas, int ln_bwd_max_ctas, int colsum_ctas, int colsum_min_rows) {
    if (ln_fwd_max_ctas > 0) g_ln_fwd_cap = ln_fwd_max_ctas;
    if (ln_bwd_max_ctas > 0) g_ln_bwd_cap = ln_bwd_max_ctas;
    if (colsum_ctas > 0) g_colsum_ctas = colsum_ctas;
    if (colsum_min_rows > 0) g_colsum_min_rows = colsum_min_rows;
    return PCM_OK;
}

// out[c] += sum over rows of src[r, c] (bias gradients); src_bf16 selects the element type.
PCM_API int pcm_colsum(long long rows, int C, const void* src, long long ld, int src_bf16, float* out,
                       pcm_stream_t stream) {
    if (rows <= 0 || C <= 0) return PCM_OK;
    if (!src || !out) return PCM_EINVAL;
    const long ctas = g_colsum_ctas > 0 ? g_colsum_ctas : 148L * 4;
    int rows_per_cta = (int)((rows + ctas - 1) / ctas);
    if (rows_per_cta < g_colsum_min_rows) rows_per_cta = g_colsum_min_rows;
    const int grid = (int)((rows + rows_per_cta - 1) / rows_per_cta);
    cudaStream_t st = pcm_cu_stream(stream);
    const int vec = src_bf16 ? 8 : 4;
    const bool vec_ok = (C % vec) == 0 && C / vec <= 256 && (ld % vec) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (vec_ok) {
        const int groups = 256 / (C / vec) > 0 ? 256 / (C / vec) : 1;
        const size_t smem = (size_t)groups * C * sizeof(float);
        if (src_bf16)
            pcm_launch(colsum_vec_kernel<__nv_bfloat16, 8>, dim3(grid), dim3(256), smem, st, reinterpret_cast<const __nv_bfloat16*>(src),
                       (long)rows, C, (long)ld, rows_per_cta, out);
        else
            pcm_launch(colsum_vec_kernel<float, 4>, dim3(grid), dim3(256), smem, st, reinterpret_cast<const float*>(src), (long)rows, C,
                       (long)ld, rows_per_cta, out);
    } else if (src_bf16) {
        pcm_launch(colsum_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(src), (long)rows, C,
                   (long)ld, rows_per_cta, out);
    } else {
        pcm_launch(colsum_kernel<float>, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float*>(src), (long)rows, C, (long)ld,
                   rows_per_cta, out);
    }
    return pcm_launch_status();
}

// n bf16 column-sum problems out_p[c] += sum_r src_p[r, c] (HOST arrays of length n; C_p % 8 == 0, C_p <= 2048, ld_p % 8 == 0,
// 16-byte aligned sources) in one launch per 64 problems.
PCM_API int pcm_colsum_grouped(int n, const void* const* src, const long long* rows, const int* C, const long long* ld,
                               float* const* out, pcm_stream_t stream) {
    if (n <= 0) return PCM_OK;
    if (!src || !rows || !C || !ld || !out) return PCM_EINVAL;
    cudaStream_t st = pcm_cu_stream(stream);
    for (int base = 0; base < n; base += CS_MAX) {
        const int cnt = n - base < CS_MAX ? n - base : CS_MAX;
        ColsumGroup g;
        g.n = cnt;
        int ctas = 0;
        size_t smem = 0;
        for (int j = 0; j < cnt; ++j) {
            const int p = base + j;
            if (!src[p] || !out[p] || rows[p] < 0) return PCM_EINVAL;
            if ((C[p] % 8) || C[p] <= 0 || C[p] / 8 > 256 || (ld[p] % 8) || (reinterpret_cast<uintptr_t>(src[p]) & 15)) return PCM_EUNSUPPORTED;
            int rpc = (int)((rows[p] + 148L * 2 - 1) / (148L * 2));
            if (rpc < 64) rpc = 64;
            g.prob[j] = ColsumProblem{reinterpret_cast<const __nv_bfloat16*>(src[p]), out[p], (long)rows[p], (long)ld[p], C[p], rpc};
            g.cta_prefix[j] = ctas;
            ctas += (int)((rows[p] + rpc - 1) / rpc);
            const int groups = 256 / (C[p] / 8) > 0 ? 256 / (C[p] / 8) : 1;
            const size_t need = (size_t)groups * C[p] * sizeof(float);
            smem = need > smem ? need : smem;
        }
        for (int j = cnt; j <= CS_MAX; ++j) g.cta_prefix[j] = ctas;
        if (ctas == 0) continue;
        cudaError_t e = pcm_launch(colsum_grouped_kernel, dim3(ctas), dim3(256), smem, st, g);
        if (e != cudaSuccess) return (int)e;
        const int r = pcm_launch_status();
        if (r) return r;
    }
    return PCM_OK;
}

// out = bf16(a + b) -- the fused `with_pos_embed` + operand cast in front of the Q/K projections
// (reference transformer.py:235-236,243).  C % 4 == 0; b may be NULL (plain cast).
PCM_API int pcm_add_cast_bf16(long long rows, int C, const float* a, const float* b, int b_row_div, void* out,
                              pcm_stream_t stream) {
    if (rows <= 0 || C <= 0) return PCM_OK;
    if (!a || !out) return PCM_EINVAL;
    if (C % 4) return PCM_EUNSUPPORTED;
    const long n4 = rows * (C / 4);
    long blocks = (n4 + 255) / 256;
    const int grid = (int)(blocks < 148L * 16 ? blocks : 148L * 16);
    pcm_launch(add_cast_bf16_kernel, dim3(grid), dim3(256), 0, pcm_cu_stream(stream), a, b, n4, C / 4, b_row_div < 1 ? 1 : b_row_div,
               reinterpret_cast<__nv_bfloat16*>(out));
    return pcm_launch_status();
}
