// attention.cu -- softmax stages of multi-head attention for sm_100a.
//
// The attention contractions (S = Q K^T, O = P V, dP = dO V^T, dQ = dS K, dK = dS^T Q,
// dV = P^T dO) run on the batched tcgen05 GEMM (pcm_gemm_bf16_ex).  These kernels are the
// row-wise stages in between; one warp owns one (batch*head, query) row:
//   forward : y = softmax(scale * s + key_padding_mask)   (fp32 scores in, bf16 probabilities out)
//             z = dropout(y)                              (counter-based hash RNG, regenerated in bwd)
//   backward: dy = dz * keep / (1 - p);  ds = scale * y * (dy - sum_j dy_j y_j)    (in place, bf16)
// Replaces the (B*h, L, S) fp32 score / softmax / dropout tensors that nn.MultiheadAttention's
// math path materialises in the reference (transformer.py:246-248; need_weights=True default).
// Buffers are [Z, Lp, Sp] with zero padding (rows L..Lp, columns S..Sp) that these kernels never
// write, so the padded K tails of the GEMMs meet exact zeros.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t dropout_bits(unsigned long long seed, unsigned long long idx) {
    // splitmix64 finaliser over (seed, element index): stateless, reproducible in backward
    unsigned long long x = seed + idx * 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return (uint32_t)(x >> 32);
}

__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(
    const float* __restrict__ S, long rows_total, int L, int Lp, int Sk, int Sp, int nh,
    const unsigned char* __restrict__ kpm, float scale, float p_drop, const unsigned long long* __restrict__ seed_base,
    unsigned long long seed_offset, __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ Zd) {
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const uint32_t thresh = (uint32_t)fminf(p_drop * 4294967296.0f, 4294967295.0f);
    const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    for (long r = wid; r < rows_total; r += nwarps) {
        const long z = r / L;
        const int l = (int)(r - z * L);
        const int b = (int)(z / nh);
        const size_t base = ((size_t)z * Lp + l) * Sp;
        const float* s = S + base;
        const unsigned char* mrow = kpm ? kpm + (size_t)b * Sk : nullptr;
        float mx = -INFINITY;
        for (int j = lane; j < Sk; j += 32) {
            const float v = (mrow && mrow[j]) ? -INFINITY : s[j] * scale;
            mx = fmaxf(mx, v);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PCM_FULL_MASK, mx, o));
        if (mx == -INFINITY) mx = 0.f;  // fully masked row (never happens on the ACT path): all zeros
        float sum = 0.f;
        for (int j = lane; j < Sk; j += 32) {
            const float v = (mrow && mrow[j]) ? -INFINITY : s[j] * scale;
            sum += __expf(v - mx);
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(PCM_FULL_MASK, sum, o);
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;
        for (int j = lane; j < Sk; j += 32) {
            const float v = (mrow && mrow[j]) ? -INFINITY : s[j] * scale;
            const float y = __expf(v - mx) * inv;
            Y[base + j] = __float2bfloat16_rn(y);
            if (Zd != Y) {
                const bool keep = dropout_bits(seed, (unsigned long long)base + j) >= thresh;
                Zd[base + j] = __float2bfloat16_rn(keep ? y * keep_scale : 0.f);
            }
        }
    }
}

__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(
    const __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ dZ, long rows_total, int L, int Lp, int Sk,
    int Sp, float scale, float p_drop, const unsigned long long* __restrict__ seed_base, unsigned long long seed_offset) {
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const uint32_t thresh = (uint32_t)fminf(p_drop * 4294967296.0f, 4294967295.0f);
    const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    for (long r = wid; r < rows_total; r += nwarps) {
        const long z = r / L;
        const int l = (int)(r - z * L);
        const size_t base = ((size_t)z * Lp + l) * Sp;
        float dot = 0.f;
        for (int j = lane; j < Sk; j += 32) {
            float dy = __bfloat162float(dZ[base + j]);
            if (p_drop > 0.f) dy = dropout_bits(seed, (unsigned long long)base + j) >= thresh ? dy * keep_scale : 0.f;
            dot = fmaf(dy, __bfloat162float(Y[base + j]), dot);
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(PCM_FULL_MASK, dot, o);
        for (int j = lane; j < Sk; j += 32) {
            float dy = __bfloat162float(dZ[base + j]);
            if (p_drop > 0.f) dy = dropout_bits(seed, (unsigned long long)base + j) >= thresh ? dy * keep_scale : 0.f;
            dZ[base + j] = __float2bfloat16_rn(scale * __bfloat162float(Y[base + j]) * (dy - dot));
        }
    }
}

inline int rows_grid(long rows) {
    long blocks = (rows + 7) / 8;  // 8 warps per 256-thread CTA
    const long cap = 148L * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

// S: fp32 [Z, Lp, Sp] raw scores (Z = B * nh); kpm: (B, Sk) bytes, non-zero = masked key (may be NULL).
// Y: bf16 softmax probabilities; Zd: bf16 dropped probabilities (pass Zd == Y when p_drop == 0).
PCM_API int pcm_attn_softmax_fwd(int Z, int L, int Lp, int Sk, int Sp, int nh, const float* S,
                                 const unsigned char* kpm, float scale, float p_drop,
                                 const unsigned long long* seed_base, unsigned long long seed_offset, void* Y,
                                 void* Zd, pcm_stream_t stream) {
    const long rows = (long)Z * L;
    if (rows <= 0 || Sk <= 0) return PCM_OK;
    if (!S || !Y || !Zd || nh <= 0 || Lp < L || Sp < Sk) return PCM_EINVAL;
    if (p_drop < 0.f || p_drop >= 1.f || (p_drop > 0.f && Zd == Y)) return PCM_EINVAL;
    attn_softmax_fwd_kernel<<<rows_grid(rows), 256, 0, pcm_cu_stream(stream)>>>(
        S, rows, L, Lp, Sk, Sp, nh, kpm, scale, p_drop, seed_base, seed_offset, reinterpret_cast<__nv_bfloat16*>(Y),
        reinterpret_cast<__nv_bfloat16*>(Zd));
    return pcm_launch_status();
}

// dZ (in) = dO V^T in bf16; dZ (out) = dS = scale * y * (dy - <dy, y>), dy = dropout-backward(dZ).
PCM_API int pcm_attn_softmax_bwd(int Z, int L, int Lp, int Sk, int Sp, const void* Y, void* dZ, float scale,
                                 float p_drop, const unsigned long long* seed_base, unsigned long long seed_offset,
                                 pcm_stream_t stream) {
    const long rows = (long)Z * L;
    if (rows <= 0 || Sk <= 0) return PCM_OK;
    if (!Y || !dZ || Lp < L || Sp < Sk || p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    attn_softmax_bwd_kernel<<<rows_grid(rows), 256, 0, pcm_cu_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(Y), reinterpret_cast<__nv_bfloat16*>(dZ), rows, L, Lp, Sk, Sp, scale,
        p_drop, seed_base, seed_offset);
    return pcm_launch_status();
}
