// unet1d.cu -- sm_100a kernels of the Diffusion-Policy denoiser:
// the non-GEMM half of `ConditionalUnet1D` (src/models/components/diffusion_policy/diffusion/
// conditional_unet1d.py:17-297, conv1d_components.py:8-45).
//
// Layout (B200-first): activations are CHANNEL-LAST (B, T, C) -- a token-major
// (B*T, C) matrix -- so that every Conv1d / ConvTranspose1d / Linear of the network is one tcgen05
// GEMM over rows = B*T and the conv weights are used IN PLACE in their torch layouts:
//   Conv1d weight (Cout, Cin, k)          viewed (Cout, Cin*k)  -- K-major B operand, column c*k+tap
//   ConvTranspose1d weight (Cin, Cout, k) viewed (Cin, Cout*k)  -- MN-major B operand
// The kernels here move data between the (B, T, C) activations and the (rows, C*k) "tap column"
// matrices those GEMMs read / write:
//   unfold: col[(b,r), c*k+tap] = x[b, r*stride + tap - pad, c]        (0 outside [0, L))
//   fold:   y[b, p, c] = bias[c] + sum_{r,tap : r*stride+tap-pad = p} col[(b,r), c*k+tap]
// Conv1d forward = unfold -> GEMM, backward dX = GEMM -> fold; ConvTranspose1d forward = GEMM ->
// fold, backward = unfold -> GEMM; weight gradients are GEMMs that land in the weight's own layout.
// HBM-bound byte movement: 2 B written per column element (unfold), 4 B read per column element
// (fold); all accesses coalesced along the column index / channel index.
//
// GroupNorm + Mish (+ FiLM scale/bias, + residual) is one kernel per direction, one CTA per
// (sample, group): the group's T x C/G tile (<= 16 KB) stays in L1 across the statistics and the
// output pass.  Statistics are two-pass (mean, then centred variance) in fp32, as torch's.
#include "common.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ float ld_any(const void* p, size_t i, int is_bf16) {
    return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                   : __ldg(reinterpret_cast<const float*>(p) + i);
}

__global__ void __launch_bounds__(256)
unfold_kernel(int B, int L, int C, int k, int stride, int pad, int R, const void* __restrict__ x, int x_bf16,
              long long ldx, __nv_bfloat16* __restrict__ col, long long ldc) {
    pcm_pdl_wait();
    const long long ncol = ldc;  // padded columns (>= C*k) are written as zeros
    const long long total = (long long)B * R * ncol;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / ncol;
        const int cc = (int)(i - row * ncol);
        float v = 0.f;
        if (cc < C * k) {
            const int c = cc / k, tap = cc - c * k;
            const int b = (int)(row / R), r = (int)(row - (long long)b * R);
            const int p = r * stride + tap - pad;
            if (p >= 0 && p < L) v = ld_any(x, ((size_t)b * L + p) * ldx + c, x_bf16);
        }
        col[i] = __float2bfloat16(v);
    }
}

__global__ void __launch_bounds__(256)
fold_kernel(int B, int L, int C, int k, int stride, int pad, int R, const float* __restrict__ col, long long ldc,
            const float* __restrict__ bias, float* __restrict__ y, __nv_bfloat16* __restrict__ yb) {
    pcm_pdl_wait();
    const long long total = (long long)B * L * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bp = i / C;
        const int p = (int)(bp % L), b = (int)(bp / L);
        float acc = bias ? __ldg(bias + c) : 0.f;
        for (int tap = 0; tap < k; ++tap) {
            const int q = p + pad - tap;
            if (q < 0 || (q % stride)) continue;
            const int r = q / stride;
            if (r < R) acc += __ldg(col + ((size_t)b * R + r) * ldc + (size_t)c * k + tap);
        }
        if (y) y[i] = acc;
        if (yb) yb[i] = __float2bfloat16(acc);
    }
}

// ---- Mish ---------------------------------------------------------------------------------------
// torch.nn.Mish: x * tanh(softplus(x)), softplus with threshold 20 (ATen Activation.cu).
__device__ __forceinline__ float softplus_f(float z) { return z > 20.f ? z : log1pf(expf(z)); }
__device__ __forceinline__ float mish_f(float z) { return z * tanhf(softplus_f(z)); }
__device__ __forceinline__ float mish_grad_f(float z) {
    const float t = tanhf(softplus_f(z));
    const float sg = 1.f / (1.f + expf(-z));
    return t + z * (1.f - t * t) * sg;
}

__device__ __forceinline__ float block_sum(float v, float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(PCM_FULL_MASK, v, o);
    __syncthreads();  // protects s_red against the previous call's readers
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += s_red[w];
    return t;
}

// One CTA per (b, g).  Thread (r, j): channels j, j+W, ... of the group, time steps r, r+rows, ...
struct GnMap {
    int W, rows, r, j;
    bool active;
    __device__ GnMap(int Cg) {
        W = Cg < (int)blockDim.x ? Cg : (int)blockDim.x;
        rows = blockDim.x / W;
        r = threadIdx.x / W;
        j = threadIdx.x - r * W;
        active = r < rows;
    }
};

__global__ void __launch_bounds__(256)
gn_mish_fwd_kernel(int T, int C, int G, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, const float* __restrict__ film,
                   const float* __restrict__ res, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    pcm_pdl_wait();
    __shared__ float s_red[32];
    const int b = blockIdx.x / G, g = blockIdx.x - b * G;
    const int Cg = C / G;
    const GnMap m(Cg);
    const size_t base = (size_t)b * T * C + (size_t)g * Cg;
    const float n = (float)T * (float)Cg;
    float s = 0.f;
    if (m.active)
        for (int t = m.r; t < T; t += m.rows)
            for (int c = m.j; c < Cg; c += m.W) s += __ldg(x + base + (size_t)t * C + c);
    const float mean = block_sum(s, s_red) / n;
    float q = 0.f;
    if (m.active)
        for (int t = m.r; t < T; t += m.rows)
            for (int c = m.j; c < Cg; c += m.W) {
                const float d = __ldg(x + base + (size_t)t * C + c) - mean;
                q += d * d;
            }
    const float rstd = rsqrtf(block_sum(q, s_red) / n + eps);
    if (threadIdx.x == 0) { mean_out[blockIdx.x] = mean; rstd_out[blockIdx.x] = rstd; }
    if (!m.active) return;
    for (int c = m.j; c < Cg; c += m.W) {
        const int ch = g * Cg + c;
        const float ga = __ldg(gamma + ch) * rstd, be = __ldg(beta + ch);
        const float fs = film ? __ldg(film + (size_t)b * 2 * C + ch) : 1.f;
        const float fb = film ? __ldg(film + (size_t)b * 2 * C + C + ch) : 0.f;
        for (int t = m.r; t < T; t += m.rows) {
            const size_t i = base + (size_t)t * C + c;
            float v = fs * mish_f((__ldg(x + i) - mean) * ga + be) + fb;
            if (res) v += __ldg(res + i);
            if (y) y[i] = v;
            if (yb) yb[i] = __float2bfloat16(v);
        }
    }
}

// Backward of y = fs * mish(gn(x)) + fb (+ res).  dres = dy is handled by the caller.
// dgamma / dbeta: one global atomic per channel per CTA; dfilm (B, 2C) is written (unique owner).
__global__ void __launch_bounds__(256)
gn_mish_bwd_kernel(int T, int C, int G, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, const float* __restrict__ film,
                   const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, float* __restrict__ dfilm) {
    pcm_pdl_wait();
    extern __shared__ float s_ch[];  // [4][Cg]: dgamma, dbeta, dscale, dbias partials
    __shared__ float s_red[32];
    const int b = blockIdx.x / G, g = blockIdx.x - b * G;
    const int Cg = C / G;
    const GnMap m(Cg);
    const size_t base = (size_t)b * T * C + (size_t)g * Cg;
    const float n = (float)T * (float)Cg;
    const float mean = __ldg(mean_in + blockIdx.x), rstd = __ldg(rstd_in + blockIdx.x);
    for (int i = threadIdx.x; i < 4 * Cg; i += blockDim.x) s_ch[i] = 0.f;
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    if (m.active)
        for (int c = m.j; c < Cg; c += m.W) {
            const int ch = g * Cg + c;
            const float ga = __ldg(gamma + ch), be = __ldg(beta + ch);
            const float fs = film ? __ldg(film + (size_t)b * 2 * C + ch) : 1.f;
            float a_dg = 0.f, a_db = 0.f, a_ds = 0.f, a_dB = 0.f;
            for (int t = m.r; t < T; t += m.rows) {
                const size_t i = base + (size_t)t * C + c;
                const float xh = (__ldg(x + i) - mean) * rstd;
                const float z = xh * ga + be;
                const float g_out = __ldg(dy + i);
                a_ds += g_out * mish_f(z);
                a_dB += g_out;
                const float dz = g_out * fs * mish_grad_f(z);
                a_dg += dz * xh;
                a_db += dz;
                const float dxh = dz * ga;
                s1 += dxh;
                s2 += dxh * xh;
            }
            if (m.rows > 1) {
                atomicAdd(&s_ch[c], a_dg); atomicAdd(&s_ch[Cg + c], a_db);
                atomicAdd(&s_ch[2 * Cg + c], a_ds); atomicAdd(&s_ch[3 * Cg + c], a_dB);
            } else {
                s_ch[c] = a_dg; s_ch[Cg + c] = a_db; s_ch[2 * Cg + c] = a_ds; s_ch[3 * Cg + c] = a_dB;
            }
        }
    s1 = block_sum(s1, s_red) / n;
    s2 = block_sum(s2, s_red) / n;  // block_sum's barriers also publish s_ch
    for (int c = threadIdx.x; c < Cg; c += blockDim.x) {
        const int ch = g * Cg + c;
        atomicAdd(dgamma + ch, s_ch[c]);
        atomicAdd(dbeta + ch, s_ch[Cg + c]);
        if (dfilm) {
            dfilm[(size_t)b * 2 * C + ch] = s_ch[2 * Cg + c];
            dfilm[(size_t)b * 2 * C + C + ch] = s_ch[3 * Cg + c];
        }
    }
    if (!m.active) return;
    for (int c = m.j; c < Cg; c += m.W) {
        const int ch = g * Cg + c;
        const float ga = __ldg(gamma + ch), be = __ldg(beta + ch);
        const float fs = film ? __ldg(film + (size_t)b * 2 * C + ch) : 1.f;
        for (int t = m.r; t < T; t += m.rows) {
            const size_t i = base + (size_t)t * C + c;
            const float xh = (__ldg(x + i) - mean) * rstd;
            const float z = xh * ga + be;
            const float dxh = __ldg(dy + i) * fs * mish_grad_f(z) * ga;
            dx[i] = rstd * (dxh - s1 - xh * s2);
        }
    }
}

// Mish on a flat fp32 vector (the shared FiLM conditioning input), forward and backward.
__global__ void __launch_bounds__(256)
mish_fwd_kernel(long long n, const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ yb) {
    pcm_pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = mish_f(__ldg(x + i));
        if (y) y[i] = v;
        if (yb) yb[i] = __float2bfloat16(v);
    }
}
__global__ void __launch_bounds__(256)
mish_bwd_kernel(long long n, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx) {
    pcm_pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = __ldg(dy + i) * mish_grad_f(__ldg(x + i));
}

inline int grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = 148LL * 16;  // 16 resident 256-thread CTAs per SM; grid-stride beyond that
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

PCM_API int pcm_conv1d_unfold(int B, int L, int C, int k, int stride, int pad, int R, const void* x, int x_bf16,
                              long long ldx, void* col, long long ldc, pcm_stream_t stream) {
    if (B <= 0 || R <= 0) return PCM_OK;
    if (!x || !col) return PCM_EINVAL;
    if (L <= 0 || C <= 0 || k <= 0 || stride <= 0 || pad < 0 || ldx < C || ldc < (long long)C * k) return PCM_EINVAL;
    const long long total = (long long)B * R * ldc;
    cudaError_t e = pcm_launch(unfold_kernel, dim3(grid_for(total)), dim3(256), 0, pcm_cu_stream(stream), B, L, C, k,
                               stride, pad, R, x, x_bf16, ldx, reinterpret_cast<__nv_bfloat16*>(col), ldc);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_conv1d_fold(int B, int L, int C, int k, int stride, int pad, int R, const float* col, long long ldc,
                            const float* bias, float* y, void* y_bf16, pcm_stream_t stream) {
    if (B <= 0 || L <= 0) return PCM_OK;
    if (!col || (!y && !y_bf16)) return PCM_EINVAL;
    if (R <= 0 || C <= 0 || k <= 0 || stride <= 0 || pad < 0 || ldc < (long long)C * k) return PCM_EINVAL;
    const long long total = (long long)B * L * C;
    cudaError_t e = pcm_launch(fold_kernel, dim3(grid_for(total)), dim3(256), 0, pcm_cu_stream(stream), B, L, C, k,
                               stride, pad, R, col, ldc, bias, y, reinterpret_cast<__nv_bfloat16*>(y_bf16));
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_groupnorm_mish_fwd(int B, int T, int C, int G, const float* x, const float* gamma, const float* beta,
                                   float eps, const float* film, const float* res, float* y, void* y_bf16,
                                   float* mean, float* rstd, pcm_stream_t stream) {
    if (B <= 0 || T <= 0) return PCM_OK;
    if (!x || !gamma || !beta || !mean || !rstd || (!y && !y_bf16)) return PCM_EINVAL;
    if (C <= 0 || G <= 0 || (C % G)) return PCM_EINVAL;
    cudaError_t e = pcm_launch(gn_mish_fwd_kernel, dim3(B * G), dim3(256), 0, pcm_cu_stream(stream), T, C, G, x, gamma,
                               beta, eps, film, res, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), mean, rstd);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_groupnorm_mish_bwd(int B, int T, int C, int G, const float* x, const float* gamma, const float* beta,
                                   const float* mean, const float* rstd, const float* film, const float* dy,
                                   float* dx, float* dgamma, float* dbeta, float* dfilm, pcm_stream_t stream) {
    if (B <= 0 || T <= 0) return PCM_OK;
    if (!x || !gamma || !beta || !mean || !rstd || !dy || !dx || !dgamma || !dbeta) return PCM_EINVAL;
    if (C <= 0 || G <= 0 || (C % G) || (film && !dfilm)) return PCM_EINVAL;
    const size_t smem = (size_t)4 * (C / G) * sizeof(float);
    if (smem > 48 * 1024) return PCM_EUNSUPPORTED;
    cudaError_t e = pcm_launch(gn_mish_bwd_kernel, dim3(B * G), dim3(256), smem, pcm_cu_stream(stream), T, C, G, x,
                               gamma, beta, mean, rstd, film, dy, dx, dgamma, dbeta, dfilm);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_mish_fwd(long long n, const float* x, float* y, void* y_bf16, pcm_stream_t stream) {
    if (n <= 0) return PCM_OK;
    if (!x || (!y && !y_bf16)) return PCM_EINVAL;
    cudaError_t e = pcm_launch(mish_fwd_kernel, dim3(grid_for(n)), dim3(256), 0, pcm_cu_stream(stream), n, x, y,
                               reinterpret_cast<__nv_bfloat16*>(y_bf16));
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

PCM_API int pcm_mish_bwd(long long n, const float* x, const float* dy, float* dx, pcm_stream_t stream) {
    if (n <= 0) return PCM_OK;
    if (!x || !dy || !dx) return PCM_EINVAL;
    cudaError_t e = pcm_launch(mish_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, pcm_cu_stream(stream), n, x, dy, dx);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}
