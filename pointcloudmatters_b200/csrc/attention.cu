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
// Buffers are [Z, Lp, Sp] with zero padding (rows L..Lp, columns S..Sp) that these kernels write
// themselves (no memsets), so the padded K tails of the GEMMs meet exact zeros.
#include "common.cuh"

namespace {

// One warp per row, the whole row in registers: lane owns float4 chunks c = lane, lane+32, ...
// (NV chunks, Sp <= 128*NV columns): one 128-bit read per chunk, one exp per element, 64-bit bf16
// stores.  Columns Sk..Sp of every row and rows L..Lp of every (batch, head) are written as zeros
// here, so the buffers never need a memset.
template <int NV>
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(
    const float* __restrict__ S, long rows_total, int L, int Lp, int Sk, int Sp, int nh,
    const unsigned char* __restrict__ kpm, float scale, float p_drop, const unsigned long long* __restrict__ seed_base,
    unsigned long long seed_offset, __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ Zd) {
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const uint32_t thr16 = pcm_drop_thr16(p_drop);
    const float keep_scale = p_drop > 0.f ? pcm_keep_scale(thr16) : 1.0f;
    const int nchunk = Sp >> 2;
    // rows_total counts PADDED rows (Z * Lp): pad rows are zero-filled
    for (long r = wid; r < rows_total; r += nwarps) {
        const long z = r / Lp;
        const int l = (int)(r - z * Lp);
        const size_t base = (size_t)r * Sp;
        if (l >= L) {
            for (int c = lane; c < nchunk; c += 32) {
                *reinterpret_cast<uint2*>(Y + base + 4 * c) = make_uint2(0u, 0u);
                if (Zd != Y) *reinterpret_cast<uint2*>(Zd + base + 4 * c) = make_uint2(0u, 0u);
            }
            continue;
        }
        const int b = (int)(z / nh);
        const unsigned char* mrow = kpm ? kpm + (size_t)b * Sk : nullptr;
        const uint32_t rseed = pcm_row_seed(seed, (unsigned long long)base);
        float v[NV][4];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (c < nchunk) t = *reinterpret_cast<const float4*>(S + base + 4 * c);
            const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * c + e;
                const bool valid = c < nchunk && j < Sk && !(mrow && mrow[j]);
                v[i][e] = valid ? tv[e] * scale : -INFINITY;
                mx = fmaxf(mx, v[i][e]);
            }
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PCM_FULL_MASK, mx, o));
        if (mx == -INFINITY) mx = 0.f;  // fully masked row (never happens on the ACT path): all zeros
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) { v[i][e] = __expf(v[i][e] - mx); sum += v[i][e]; }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(PCM_FULL_MASK, sum, o);
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c >= nchunk) continue;
            float y[4], zd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { y[e] = v[i][e] * inv; zd[e] = y[e]; }  // exactly 0 for masked / padded columns
            if (Zd != Y) {
                const uint32_t h0 = pcm_pair_bits(rseed, 2 * c), h1 = pcm_pair_bits(rseed, 2 * c + 1);
                zd[0] = (h0 & 0xFFFFu) >= thr16 ? y[0] * keep_scale : 0.f;
                zd[1] = (h0 >> 16) >= thr16 ? y[1] * keep_scale : 0.f;
                zd[2] = (h1 & 0xFFFFu) >= thr16 ? y[2] * keep_scale : 0.f;
                zd[3] = (h1 >> 16) >= thr16 ? y[3] * keep_scale : 0.f;
            }
            __nv_bfloat162 a0 = __floats2bfloat162_rn(y[0], y[1]), a1 = __floats2bfloat162_rn(y[2], y[3]);
            *reinterpret_cast<uint2*>(Y + base + 4 * c) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
            if (Zd != Y) {
                __nv_bfloat162 b0 = __floats2bfloat162_rn(zd[0], zd[1]), b1 = __floats2bfloat162_rn(zd[2], zd[3]);
                *reinterpret_cast<uint2*>(Zd + base + 4 * c) = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
            }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(
    const __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ dZ, long rows_total, int L, int Lp, int Sk,
    int Sp, float scale, float p_drop, const unsigned long long* __restrict__ seed_base, unsigned long long seed_offset) {
    const unsigned long long seed = (seed_base ? *seed_base : 0ULL) * 0xD1342543DE82EF95ULL + seed_offset;
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const uint32_t thr16 = pcm_drop_thr16(p_drop);
    const float keep_scale = p_drop > 0.f ? pcm_keep_scale(thr16) : 1.0f;
    const int nchunk = Sp >> 2;
    for (long r = wid; r < rows_total; r += nwarps) {
        const long z = r / Lp;
        const int l = (int)(r - z * Lp);
        const size_t base = (size_t)r * Sp;
        const uint32_t rseed = pcm_row_seed(seed, (unsigned long long)base);
        if (l >= L) {  // pad rows of dS must be exact zeros (K tail of the dK GEMM)
            for (int c = lane; c < nchunk; c += 32) *reinterpret_cast<uint2*>(dZ + base + 4 * c) = make_uint2(0u, 0u);
            continue;
        }
        float y[NV][4], dy[NV][4];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            uint2 yr = make_uint2(0u, 0u), dr = make_uint2(0u, 0u);
            if (c < nchunk) {
                yr = *reinterpret_cast<const uint2*>(Y + base + 4 * c);
                dr = *reinterpret_cast<const uint2*>(dZ + base + 4 * c);
            }
            const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&yr);
            const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&dr);
            const float2 y01 = __bfloat1622float2(yp[0]), y23 = __bfloat1622float2(yp[1]);
            const float2 d01 = __bfloat1622float2(dp[0]), d23 = __bfloat1622float2(dp[1]);
            y[i][0] = y01.x; y[i][1] = y01.y; y[i][2] = y23.x; y[i][3] = y23.y;
            dy[i][0] = d01.x; dy[i][1] = d01.y; dy[i][2] = d23.x; dy[i][3] = d23.y;
            if (p_drop > 0.f) {
                const uint32_t h0 = pcm_pair_bits(rseed, 2 * c), h1 = pcm_pair_bits(rseed, 2 * c + 1);
                dy[i][0] = (h0 & 0xFFFFu) >= thr16 ? dy[i][0] * keep_scale : 0.f;
                dy[i][1] = (h0 >> 16) >= thr16 ? dy[i][1] * keep_scale : 0.f;
                dy[i][2] = (h1 & 0xFFFFu) >= thr16 ? dy[i][2] * keep_scale : 0.f;
                dy[i][3] = (h1 >> 16) >= thr16 ? dy[i][3] * keep_scale : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * c + e;
                if (c >= nchunk || j >= Sk) { dy[i][e] = 0.f; y[i][e] = 0.f; }
                dot = fmaf(dy[i][e], y[i][e], dot);
            }
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(PCM_FULL_MASK, dot, o);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c >= nchunk) continue;
            float ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) ds[e] = scale * y[i][e] * (dy[i][e] - dot);  // 0 for padded columns (y = 0)
            __nv_bfloat162 a0 = __floats2bfloat162_rn(ds[0], ds[1]), a1 = __floats2bfloat162_rn(ds[2], ds[3]);
            *reinterpret_cast<uint2*>(dZ + base + 4 * c) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
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
    const long rows = (long)Z * Lp;  // padded rows are zero-filled by the kernel
    if (rows <= 0 || Sk <= 0) return PCM_OK;
    if (!S || !Y || !Zd || nh <= 0 || Lp < L || Sp < Sk) return PCM_EINVAL;
    if (p_drop < 0.f || p_drop >= 1.f || (p_drop > 0.f && Zd == Y)) return PCM_EINVAL;
    if (Sp % 4 || Sp > 4096) return PCM_EUNSUPPORTED;
    cudaStream_t st = pcm_cu_stream(stream);
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(Y);
    __nv_bfloat16* zd = reinterpret_cast<__nv_bfloat16*>(Zd);
    const int g = rows_grid(rows);
#define FWD(NV) attn_softmax_fwd_kernel<NV><<<g, 256, 0, st>>>(S, rows, L, Lp, Sk, Sp, nh, kpm, scale, p_drop, seed_base, seed_offset, y, zd)
    if (Sp <= 128) FWD(1); else if (Sp <= 256) FWD(2); else if (Sp <= 512) FWD(4); else if (Sp <= 1024) FWD(8);
    else if (Sp <= 2048) FWD(16); else FWD(32);
#undef FWD
    return pcm_launch_status();
}

// dZ (in) = dO V^T in bf16; dZ (out) = dS = scale * y * (dy - <dy, y>), dy = dropout-backward(dZ).
PCM_API int pcm_attn_softmax_bwd(int Z, int L, int Lp, int Sk, int Sp, const void* Y, void* dZ, float scale,
                                 float p_drop, const unsigned long long* seed_base, unsigned long long seed_offset,
                                 pcm_stream_t stream) {
    const long rows = (long)Z * Lp;
    if (rows <= 0 || Sk <= 0) return PCM_OK;
    if (!Y || !dZ || Lp < L || Sp < Sk || p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    if (Sp % 4 || Sp > 4096) return PCM_EUNSUPPORTED;
    cudaStream_t st = pcm_cu_stream(stream);
    const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(Y);
    __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(dZ);
    const int g = rows_grid(rows);
#define BWD(NV) attn_softmax_bwd_kernel<NV><<<g, 256, 0, st>>>(y, dz, rows, L, Lp, Sk, Sp, scale, p_drop, seed_base, seed_offset)
    if (Sp <= 128) BWD(1); else if (Sp <= 256) BWD(2); else if (Sp <= 512) BWD(4); else if (Sp <= 1024) BWD(8);
    else if (Sp <= 2048) BWD(16); else BWD(32);
#undef BWD
    return pcm_launch_status();
}
