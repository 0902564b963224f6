// flash_attn.cu -- fused multi-head attention (head_dim 64) forward and backward on tcgen05.
//
//     O = dropout(softmax(scale * Q K^T + key_padding_mask)) V          per (batch, head)
//
// Replaces nn.MultiheadAttention's math path of the reference (transformer.py:246-248,329-340),
// which materialises the (B*h, L, S) score / probability / dropout tensors in HBM four times per
// layer and direction.  Here a score tile never leaves the SM:
//
//   forward, one CTA per (batch*head, 128-query tile), 2 CTAs resident per SM:
//     warp 0   TMA producer   Q tile once, K tiles through a 2-stage ring, V tiles (1 stage)
//     warp 1   MMA issuer     S_j = Q K_j^T   (UMMA 128 x nkv x 16, fp32 accumulator in TMEM)
//                             PV_j = P_j V_j  (UMMA 128 x 64 x 16, P_j from shared memory)
//     warps 2-5 softmax       one thread per query row (TMEM lane): tcgen05.ld the row of S_j,
//                             online max / exp2 / row sum, counter-based dropout, bf16 P_j into the
//                             canonical SWIZZLE_128B K-major tile; the running output row (64 fp32)
//                             lives in registers:  O <- alpha * (O + PV_{j-1})
//   backward, one CTA per (batch*head, 128-key tile), looping over query tiles (FlashAttention-2
//   schedule): S = Q K^T and dP = dO V^T in TMEM -> 8 softmax warps rebuild P from the saved
//   log-sum-exp, form dS = scale * P o (dropout'(dP) - delta) and write bf16 dropout(P) and dS
//   tiles to shared memory ONCE; the same bytes serve as the MN-major A operand of
//   dV += P^T dO, dK += dS^T Q and as the K-major A operand of dQ_i = dS K.  dK / dV accumulate in
//   TMEM over the query loop; dQ tiles are reduced across key tiles with 16-byte red.global.add.
//
// Layouts: Q, K, V, dO are "head-split" bf16 (B*nh, rows, 64) (the in-projection GEMM's epilogue
// writes them that way); O, dQ, dK, dV are token-major bf16 (row = l * B + b, column = h * 64 + d),
// what the out-projection / in-projection-gradient GEMMs read.  Operand tiles are fetched with
// rank-3 tensor maps (d, row, batch*head) so rows past L / S arrive as zeros.
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace {

using namespace pcm_tc;

constexpr uint32_t TILE_BYTES = 128 * 64 * 2;  // one [128 rows x 64] bf16 operand tile = 16 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr int MAX_KEYS = 8192;

struct FlashParams {
    int B, nh, L, S;
    const unsigned char* kpm;  // (B, S) bytes, non-zero = masked key; may be NULL
    float scale, scale_log2;   // softmax scale, and scale * log2(e)
    uint32_t thr16;            // dropout threshold on 16 random bits (0 = no dropout)
    float keep_scale;
    const unsigned long long* seed_base;
    unsigned long long seed_offset;
    // forward outputs
    __nv_bfloat16* O;
    int ldo;
    float* lse;  // (B*nh, L): log2-domain log-sum-exp of the scaled scores
    // backward
    const float* delta;  // (B*nh, L): rowsum(dO o O)
    float* dQacc;        // (B*nh, L, 64) fp32, zero-initialised
    __nv_bfloat16* dK;
    __nv_bfloat16* dV;
    int ldkv;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// 16-byte piece `piece` (0..7) of row `row` of 64-column block `blk` in a [128 x 128] bf16 tile
// stored as two SWIZZLE_128B [128 rows x 128 B] blocks (what a TMA box with that swizzle writes).
__device__ __forceinline__ uint32_t sw128_off(int blk, int row, int piece) {
    return (uint32_t)(blk * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((piece ^ (row & 7)) << 4));
}

__device__ __forceinline__ void st_shared_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

// valid-key bit mask of the CTA's key range: word w covers keys key0 + 32 w .. + 31
__device__ __forceinline__ void build_key_bits(uint32_t* kbits, int nwords, int key0, const FlashParams& p, int b,
                                               int warp, int nwarps, int lane) {
    for (int w = warp; w < nwords; w += nwarps) {
        const int col = key0 + w * 32 + lane;
        const bool valid = col < p.S && !(p.kpm != nullptr && p.kpm[(size_t)b * p.S + col] != 0);
        const uint32_t bits = __ballot_sync(PCM_FULL_MASK, valid);
        if (lane == 0) kbits[w] = bits;
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 192;
constexpr uint32_t FWD_SMEM_TILES = 6 * TILE_BYTES;  // Q, K0, K1, V, P (2 blocks)
constexpr uint32_t FWD_SMEM = FWD_SMEM_TILES + 256 + MAX_KEYS / 8 + 1024;

__global__ void __launch_bounds__(FWD_THREADS, 2) flash_fwd_kernel(const __grid_constant__ CUtensorMap tq,
                                                                    const __grid_constant__ CUtensorMap tk,
                                                                    const __grid_constant__ CUtensorMap tv,
                                                                    const FlashParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = sm;
    uint8_t* sK = sm + TILE_BYTES;      // 2 stages
    uint8_t* sV = sm + 3 * TILE_BYTES;  // 1 stage
    uint8_t* sP = sm + 4 * TILE_BYTES;  // [128 x 128] bf16
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + FWD_SMEM_TILES);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* v_full = bars + 5;
    uint64_t* v_empty = bars + 6;
    uint64_t* s_full = bars + 7;
    uint64_t* p_full = bars + 8;
    uint64_t* pv_full = bars + 9;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
    uint32_t* kbits = reinterpret_cast<uint32_t*>(sm + FWD_SMEM_TILES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq_tiles = (p.L + 127) >> 7;
    const int z = blockIdx.x / nq_tiles;
    const int q0 = (blockIdx.x - z * nq_tiles) << 7;
    const int b = z / p.nh, h = z - b * p.nh;
    const int n_kv = (p.S + 127) >> 7;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 256);
    build_key_bits(kbits, n_kv * 4, 0, p, b, warp, FWD_THREADS / 32, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S = 0, TM_PV = 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(sQ, &tq, q_full, 0, q0, z);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                mbar_wait(&k_empty[st], (((uint32_t)j >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[st], TILE_BYTES);
                tma_load_3d(sK + st * TILE_BYTES, &tk, &k_full[st], 0, j << 7, z);
                mbar_wait(v_empty, ((uint32_t)j & 1) ^ 1);
                mbar_expect_tx(v_full, TILE_BYTES);
                tma_load_3d(sV, &tv, v_full, 0, j << 7, z);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
            auto issue_s = [&](int j) {
                const int st = j & 1;
                const int nkv16 = (min(128, p.S - (j << 7)) + 15) & ~15;
                mbar_wait(&k_full[st], ((uint32_t)j >> 1) & 1);
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16(128, nkv16, false, false);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + TM_S, make_smem_desc(aQ + k * 32, 16, 1024),
                             make_smem_desc(aK + st * TILE_BYTES + k * 32, 16, 1024), idesc, k != 0 ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(&k_empty[st]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            const uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(p_full, (uint32_t)j & 1);  // P_j in smem; S_j and PV_{j-1} have been read out of TMEM
                tc_fence_after();
                if (j + 1 < n_kv) issue_s(j + 1);
                mbar_wait(v_full, (uint32_t)j & 1);
                tc_fence_after();
                const int ksteps = ((min(128, p.S - (j << 7)) + 15) & ~15) >> 4;
                for (int ks = 0; ks < ksteps; ++ks)
                    umma_f16(tmem_base + TM_PV, make_smem_desc(aP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                             make_smem_desc(aV + ks * 2048, 16384, 1024), idesc_pv, ks != 0 ? 1u : 0u);
                umma_commit(pv_full);
                umma_commit(v_empty);
            }
        }
    } else {
        // ===== softmax warps: thread = query row =====
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int l = q0 + row;
        const bool warp_active = q0 + quad * 32 < p.L;  // warp-uniform
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        const unsigned long long seed = (p.seed_base ? *p.seed_base : 0ULL) * 0xD1342543DE82EF95ULL + p.seed_offset;
        const uint32_t rseed = pcm_row_seed(seed, (unsigned long long)z * p.L + l);
        const uint32_t thr16 = p.thr16;
        float o[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) o[e] = 0.f;
        float m = -INFINITY, lsum = 0.f;
        uint32_t v[32];
        for (int j = 0; j < n_kv; ++j) {
            const int nkv = min(128, p.S - (j << 7));
            mbar_wait(s_full, (uint32_t)j & 1);
            tc_fence_after();
            float alpha = 1.f, m_use = 0.f, m_new = m;
            if (warp_active) {
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    if (c * 32 >= nkv) break;
                    tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                    tmem_ld_wait(v);
                    const uint32_t bits = kbits[j * 4 + c];
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        mx = fmaxf(mx, ((bits >> e) & 1u) ? __uint_as_float(v[e]) * p.scale_log2 : -INFINITY);
                }
                m_new = fmaxf(m, mx);
                m_use = m_new == -INFINITY ? 0.f : m_new;
                alpha = fast_exp2(m - m_use);
            }
            if (j > 0) {
                mbar_wait(pv_full, ((uint32_t)j - 1) & 1);
                tc_fence_after();
                if (warp_active) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tmem_ld_32x32b_x32(t_row + TM_PV + c * 32, v);
                        tmem_ld_wait(v);
#pragma unroll
                        for (int e = 0; e < 32; ++e) o[c * 32 + e] += __uint_as_float(v[e]);
                    }
                }
            }
            if (warp_active) {
#pragma unroll
                for (int e = 0; e < 64; ++e) o[e] *= alpha;
                lsum *= alpha;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    if (c * 32 >= nkv) break;
                    tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                    tmem_ld_wait(v);
                    const uint32_t bits = kbits[j * 4 + c];
                    const uint32_t pair0 = (uint32_t)((j << 7) + c * 32) >> 1;
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float p0 = ((bits >> e) & 1u) ? fast_exp2(__uint_as_float(v[e]) * p.scale_log2 - m_use) : 0.f;
                        float p1 = ((bits >> (e + 1)) & 1u) ? fast_exp2(__uint_as_float(v[e + 1]) * p.scale_log2 - m_use) : 0.f;
                        lsum += p0 + p1;
                        if (thr16 != 0) {
                            const uint32_t hb = pcm_pair_bits(rseed, pair0 + (e >> 1));
                            p0 = (hb & 0xFFFFu) >= thr16 ? p0 : 0.f;
                            p1 = (hb >> 16) >= thr16 ? p1 : 0.f;
                        }
                        pk[e >> 1] = pack_bf16(p0, p1);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        st_shared_v4(sP + sw128_off(c >> 1, row, (c & 1) * 4 + i), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2],
                                     pk[4 * i + 3]);
                }
                m = m_new;
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_full, ((uint32_t)n_kv - 1) & 1);
        tc_fence_after();
        if (warp_active) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tmem_ld_32x32b_x32(t_row + TM_PV + c * 32, v);
                tmem_ld_wait(v);
#pragma unroll
                for (int e = 0; e < 32; ++e) o[c * 32 + e] += __uint_as_float(v[e]);
            }
            if (l < p.L) {
                const float inv = lsum > 0.f ? p.keep_scale / lsum : 0.f;
                uint4* dst = reinterpret_cast<uint4*>(p.O + ((size_t)l * p.B + b) * p.ldo + h * 64);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = make_uint4(pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                                        pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
                p.lse[(size_t)z * p.L + l] = lsum > 0.f ? m + log2f(lsum) : INFINITY;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// delta[z, l] = sum_d dO[z, l, d] * O[l * B + b, h * 64 + d]      (one warp per row)
__global__ void __launch_bounds__(256) flash_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                                          int ldo, int B, int nh, int L, long rows, float* __restrict__ delta) {
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = wid; r < rows; r += nwarps) {
        const int z = (int)(r / L), l = (int)(r - (long)z * L);
        const int b = z / nh, h = z - b * nh;
        const __nv_bfloat162 a = reinterpret_cast<const __nv_bfloat162*>(dO + (size_t)r * 64)[lane];
        const __nv_bfloat162 c = reinterpret_cast<const __nv_bfloat162*>(O + ((size_t)l * B + b) * ldo + h * 64)[lane];
        const float2 af = __bfloat1622float2(a), cf = __bfloat1622float2(c);
        float s = af.x * cf.x + af.y * cf.y;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(PCM_FULL_MASK, s, o);
        if (lane == 0) delta[r] = s;
    }
}

// dQ token-major bf16 <- fp32 (Z, L, 64) accumulator          (one warp per row, 8 bytes per lane)
__global__ void __launch_bounds__(256) flash_dq_store_kernel(const float* __restrict__ acc, int B, int nh, int L, long rows,
                                                             __nv_bfloat16* __restrict__ dQ, int ldq) {
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = wid; r < rows; r += nwarps) {
        const int z = (int)(r / L), l = (int)(r - (long)z * L);
        const int b = z / nh, h = z - b * nh;
        const float2 a = reinterpret_cast<const float2*>(acc + (size_t)r * 64)[lane];
        reinterpret_cast<uint32_t*>(dQ + ((size_t)l * B + b) * ldq + h * 64)[lane] = pack_bf16(a.x, a.y);
    }
}

constexpr int BWD_THREADS = 320;
constexpr uint32_t BWD_SMEM_TILES = 10 * TILE_BYTES;  // K, V, Q[2], dO[2], Pd (2 blocks), dS (2 blocks)
constexpr uint32_t BWD_SMEM = BWD_SMEM_TILES + 256 + 1024;

__global__ void __launch_bounds__(BWD_THREADS, 1) flash_bwd_kernel(const __grid_constant__ CUtensorMap tq,
                                                                    const __grid_constant__ CUtensorMap tk,
                                                                    const __grid_constant__ CUtensorMap tv,
                                                                    const __grid_constant__ CUtensorMap tdo,
                                                                    const FlashParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sK = sm;
    uint8_t* sV = sm + TILE_BYTES;
    uint8_t* sQ = sm + 2 * TILE_BYTES;   // 2 stages
    uint8_t* sdO = sm + 4 * TILE_BYTES;  // 2 stages
    uint8_t* sPd = sm + 6 * TILE_BYTES;  // [128 q x 128 kv] bf16
    uint8_t* sdS = sm + 8 * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BWD_SMEM_TILES);
    uint64_t* kv_full = bars + 0;
    uint64_t* qdo_full = bars + 1;   // [2]
    uint64_t* qdo_empty = bars + 3;  // [2]
    uint64_t* sdp_full = bars + 5;
    uint64_t* pds_full = bars + 6;
    uint64_t* dq_full = bars + 7;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
    uint32_t* kbits = reinterpret_cast<uint32_t*>(bars + 20);  // 4 words

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kv = (p.S + 127) >> 7;
    const int z = blockIdx.x / n_kv;
    const int jt = blockIdx.x - z * n_kv;
    const int kv0 = jt << 7;
    const int b = z / p.nh, h = z - b * p.nh;
    const int nq_tiles = (p.L + 127) >> 7;
    const int nkv = min(128, p.S - kv0);
    const int nkv16 = (nkv + 15) & ~15;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv); prefetch_tmap(&tdo);
        mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
        mbar_init(sdp_full, 1); mbar_init(pds_full, 8); mbar_init(dq_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
    build_key_bits(kbits, 4, kv0, p, b, warp, BWD_THREADS / 32, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * TILE_BYTES);
            tma_load_3d(sK, &tk, kv_full, 0, kv0, z);
            tma_load_3d(sV, &tv, kv_full, 0, kv0, z);
            for (int i = 0; i < nq_tiles; ++i) {
                const int st = i & 1;
                mbar_wait(&qdo_empty[st], (((uint32_t)i >> 1) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[st], 2 * TILE_BYTES);
                tma_load_3d(sQ + st * TILE_BYTES, &tq, &qdo_full[st], 0, i << 7, z);
                tma_load_3d(sdO + st * TILE_BYTES, &tdo, &qdo_full[st], 0, i << 7, z);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO);
            const uint32_t aPd = smem_u32(sPd), adS = smem_u32(sdS);
            const uint32_t idesc_s = make_idesc_bf16(128, nkv16, false, false);
            const uint32_t idesc_t = make_idesc_bf16(128, 64, true, true);    // dV, dK: A and B MN-major
            const uint32_t idesc_dq = make_idesc_bf16(128, 64, false, true);  // dQ: A K-major, B MN-major
            auto issue_sdp = [&](int i) {
                const int st = i & 1;
                mbar_wait(&qdo_full[st], ((uint32_t)i >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + TM_S, make_smem_desc(aQ + st * TILE_BYTES + k * 32, 16, 1024),
                             make_smem_desc(aK + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + TM_DP, make_smem_desc(adO + st * TILE_BYTES + k * 32, 16, 1024),
                             make_smem_desc(aV + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(sdp_full);
            };
            mbar_wait(kv_full, 0);
            issue_sdp(0);
            for (int i = 0; i < nq_tiles; ++i) {
                const int st = i & 1;
                mbar_wait(pds_full, (uint32_t)i & 1);  // Pd_i, dS_i in smem; S_i, dP_i, dQ_{i-1} read out of TMEM
                tc_fence_after();
                if (i + 1 < nq_tiles) issue_sdp(i + 1);
                const int qsteps = ((min(128, p.L - (i << 7)) + 15) & ~15) >> 4;  // query rows are the K dimension
                for (int ks = 0; ks < qsteps; ++ks)
                    umma_f16(tmem_base + TM_DV, make_smem_desc(aPd + ks * 2048, 16384, 1024),
                             make_smem_desc(adO + st * TILE_BYTES + ks * 2048, 16384, 1024), idesc_t, (i | ks) != 0 ? 1u : 0u);
                for (int ks = 0; ks < qsteps; ++ks)
                    umma_f16(tmem_base + TM_DK, make_smem_desc(adS + ks * 2048, 16384, 1024),
                             make_smem_desc(aQ + st * TILE_BYTES + ks * 2048, 16384, 1024), idesc_t, (i | ks) != 0 ? 1u : 0u);
                const int ksteps = nkv16 >> 4;
                for (int ks = 0; ks < ksteps; ++ks)
                    umma_f16(tmem_base + TM_DQ, make_smem_desc(adS + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                             make_smem_desc(aK + ks * 2048, 16384, 1024), idesc_dq, ks != 0 ? 1u : 0u);
                umma_commit(dq_full);
                umma_commit(&qdo_empty[st]);
            }
        }
    } else {
        // ===== softmax / gradient warps: thread = (query row, 64-key half) =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        const unsigned long long seed = (p.seed_base ? *p.seed_base : 0ULL) * 0xD1342543DE82EF95ULL + p.seed_offset;
        const uint32_t thr16 = p.thr16;
        const float keep_scale = p.keep_scale;
        uint32_t v[32], w[32];

        auto flush_dq = [&](int i_prev) {
            // dQ_{i_prev} columns [half*32, +32) of this thread's row -> global fp32 accumulator
            tmem_ld_32x32b_x32(t_row + TM_DQ + half * 32, v);
            tmem_ld_wait(v);
        };
        auto red_dq = [&](int i_prev) {
            const int lq = (i_prev << 7) + row;
            if (lq < p.L) {
                float* dst = p.dQacc + ((size_t)z * p.L + lq) * 64 + half * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                    red_add_v4(dst + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                               __uint_as_float(v[e + 3]));
            }
        };

        for (int i = 0; i < nq_tiles; ++i) {
            const int nq = min(128, p.L - (i << 7));
            const int nq16 = (nq + 15) & ~15;
            const bool warp_active = quad * 32 < nq16;  // warp-uniform: rows this warp owns are read by the MMAs
            const int l = (i << 7) + row;
            const bool row_valid = row < nq;
            const float lse_r = row_valid ? p.lse[(size_t)z * p.L + l] : INFINITY;
            const float delta_r = row_valid ? p.delta[(size_t)z * p.L + l] : 0.f;
            const uint32_t rseed = pcm_row_seed(seed, (unsigned long long)z * p.L + l);
            uint32_t pd_pk[32], ds_pk[32];
            mbar_wait(sdp_full, (uint32_t)i & 1);
            tc_fence_after();
            if (warp_active) {
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int c = half * 2 + c2;
                    if (c * 32 < nkv16) {
                        tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                        tmem_ld_32x32b_x32(t_row + TM_DP + c * 32, w);
                        tmem_ld_wait(v);
                        tmem_ld_wait(w);
                        const uint32_t bits = row_valid ? kbits[c] : 0u;
                        const uint32_t pair0 = (uint32_t)(kv0 + c * 32) >> 1;
#pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            const float p0 = ((bits >> e) & 1u) ? fast_exp2(__uint_as_float(v[e]) * p.scale_log2 - lse_r) : 0.f;
                            const float p1 = ((bits >> (e + 1)) & 1u) ? fast_exp2(__uint_as_float(v[e + 1]) * p.scale_log2 - lse_r) : 0.f;
                            float pd0 = p0, pd1 = p1;
                            float g0 = __uint_as_float(w[e]), g1 = __uint_as_float(w[e + 1]);
                            if (thr16 != 0) {
                                const uint32_t hb = pcm_pair_bits(rseed, pair0 + (e >> 1));
                                const bool k0 = (hb & 0xFFFFu) >= thr16, k1 = (hb >> 16) >= thr16;
                                pd0 = k0 ? p0 * keep_scale : 0.f; g0 = k0 ? g0 * keep_scale : 0.f;
                                pd1 = k1 ? p1 * keep_scale : 0.f; g1 = k1 ? g1 * keep_scale : 0.f;
                            }
                            // select (not multiply by zero): columns past nkv16 hold uninitialised TMEM
                            const float s0 = ((bits >> e) & 1u) ? p0 * (g0 - delta_r) * p.scale : 0.f;
                            const float s1 = ((bits >> (e + 1)) & 1u) ? p1 * (g1 - delta_r) * p.scale : 0.f;
                            pd_pk[c2 * 16 + (e >> 1)] = pack_bf16(pd0, pd1);
                            ds_pk[c2 * 16 + (e >> 1)] = pack_bf16(s0, s1);
                        }
                    }
                }
            }
            if (i > 0) {
                mbar_wait(dq_full, ((uint32_t)i - 1) & 1);  // tile i-1's MMAs are done: Pd / dS smem free, dQ_{i-1} ready
                tc_fence_after();
            }
            if (warp_active) {
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int c = half * 2 + c2;
                    if (c * 32 < nkv16) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t off = sw128_off(c >> 1, row, (c & 1) * 4 + q);
                            st_shared_v4(sPd + off, pd_pk[c2 * 16 + 4 * q], pd_pk[c2 * 16 + 4 * q + 1], pd_pk[c2 * 16 + 4 * q + 2],
                                         pd_pk[c2 * 16 + 4 * q + 3]);
                            st_shared_v4(sdS + off, ds_pk[c2 * 16 + 4 * q], ds_pk[c2 * 16 + 4 * q + 1], ds_pk[c2 * 16 + 4 * q + 2],
                                         ds_pk[c2 * 16 + 4 * q + 3]);
                        }
                    }
                }
            }
            fence_proxy_async();
            if (i > 0) flush_dq(i - 1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full);
            if (i > 0) red_dq(i - 1);
        }
        mbar_wait(dq_full, ((uint32_t)nq_tiles - 1) & 1);
        tc_fence_after();
        flush_dq(nq_tiles - 1);
        red_dq(nq_tiles - 1);
        // dV, dK: this thread's key row, columns [half*32, +32)
        const int s_row = kv0 + row;
        tmem_ld_32x32b_x32(t_row + TM_DV + half * 32, v);
        tmem_ld_32x32b_x32(t_row + TM_DK + half * 32, w);
        tmem_ld_wait(v);
        tmem_ld_wait(w);
        if (s_row < p.S) {
            const size_t off = ((size_t)s_row * p.B + b) * p.ldkv + h * 64 + half * 32;
            uint4* dv = reinterpret_cast<uint4*>(p.dV + off);
            uint4* dk = reinterpret_cast<uint4*>(p.dK + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                dv[q] = make_uint4(pack_bf16(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1])),
                                   pack_bf16(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                                   pack_bf16(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                                   pack_bf16(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
                dk[q] = make_uint4(pack_bf16(__uint_as_float(w[8 * q]), __uint_as_float(w[8 * q + 1])),
                                   pack_bf16(__uint_as_float(w[8 * q + 2]), __uint_as_float(w[8 * q + 3])),
                                   pack_bf16(__uint_as_float(w[8 * q + 4]), __uint_as_float(w[8 * q + 5])),
                                   pack_bf16(__uint_as_float(w[8 * q + 6]), __uint_as_float(w[8 * q + 7])));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int head_split_map(const void* ptr, int Z, int rows, CUtensorMap* out) {
    const uint64_t dims[3] = {64, (uint64_t)rows, (uint64_t)Z};
    const uint64_t strides[2] = {128, (uint64_t)rows * 128};
    const uint32_t box[3] = {64, 128, 1};
    return tensor_map_bf16(ptr, 3, dims, strides, box, out);
}

int fill_common(FlashParams& p, int B, int nh, int L, int S, const unsigned char* kpm, float scale, float p_drop,
                const unsigned long long* seed_base, unsigned long long seed_offset) {
    if (B <= 0 || nh <= 0 || L <= 0 || S <= 0) return PCM_EINVAL;
    if (p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    if (S > MAX_KEYS) return PCM_EUNSUPPORTED;
    p.B = B; p.nh = nh; p.L = L; p.S = S; p.kpm = kpm;
    p.scale = scale; p.scale_log2 = scale * LOG2E;
    p.thr16 = p_drop > 0.f ? pcm_drop_thr16(p_drop) : 0u;
    p.keep_scale = p.thr16 ? pcm_keep_scale(p.thr16) : 1.0f;
    p.seed_base = seed_base; p.seed_offset = seed_offset;
    return PCM_OK;
}

inline bool misaligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) != 0; }

}  // namespace

PCM_API int pcm_flash_attn_fwd(int B, int nh, int L, int S, const void* Q, const void* K, const void* V,
                               const unsigned char* kpm, float scale, float p_drop,
                               const unsigned long long* seed_base, unsigned long long seed_offset, void* O, int ldo,
                               float* lse, pcm_stream_t stream) {
    if (!Q || !K || !V || !O || !lse) return PCM_EINVAL;
    if (misaligned16(Q) || misaligned16(K) || misaligned16(V) || misaligned16(O) || (ldo % 8)) return PCM_EUNSUPPORTED;
    FlashParams p{};
    int r = fill_common(p, B, nh, L, S, kpm, scale, p_drop, seed_base, seed_offset);
    if (r) return r;
    p.O = reinterpret_cast<__nv_bfloat16*>(O); p.ldo = ldo; p.lse = lse;
    const int Z = B * nh;
    CUtensorMap tq, tk, tv;
    if ((r = head_split_map(Q, Z, L, &tq))) return r;
    if ((r = head_split_map(K, Z, S, &tk))) return r;
    if ((r = head_split_map(V, Z, S, &tv))) return r;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const long grid = (long)Z * ((L + 127) / 128);
    flash_fwd_kernel<<<(unsigned)grid, FWD_THREADS, FWD_SMEM, pcm_cu_stream(stream)>>>(tq, tk, tv, p);
    return pcm_launch_status();
}

PCM_API int pcm_flash_attn_bwd(int B, int nh, int L, int S, const void* Q, const void* K, const void* V, const void* O,
                               int ldo, const void* dO, const float* lse, const unsigned char* kpm, float scale,
                               float p_drop, const unsigned long long* seed_base, unsigned long long seed_offset,
                               float* delta, float* dQacc, void* dQ, int ldq, void* dK, void* dV, int ldkv,
                               pcm_stream_t stream) {
    if (!Q || !K || !V || !O || !dO || !lse || !delta || !dQacc || !dQ || !dK || !dV) return PCM_EINVAL;
    if (misaligned16(Q) || misaligned16(K) || misaligned16(V) || misaligned16(dO) || misaligned16(dK) || misaligned16(dV) ||
        (ldkv % 8) || (ldq % 2) || (ldo % 2))
        return PCM_EUNSUPPORTED;
    FlashParams p{};
    int r = fill_common(p, B, nh, L, S, kpm, scale, p_drop, seed_base, seed_offset);
    if (r) return r;
    p.lse = const_cast<float*>(lse); p.delta = delta; p.dQacc = dQacc;
    p.dK = reinterpret_cast<__nv_bfloat16*>(dK); p.dV = reinterpret_cast<__nv_bfloat16*>(dV); p.ldkv = ldkv;
    const int Z = B * nh;
    CUtensorMap tq, tk, tv, tdo;
    if ((r = head_split_map(Q, Z, L, &tq))) return r;
    if ((r = head_split_map(K, Z, S, &tk))) return r;
    if ((r = head_split_map(V, Z, S, &tv))) return r;
    if ((r = head_split_map(dO, Z, L, &tdo))) return r;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    cudaStream_t st = pcm_cu_stream(stream);
    const long rows = (long)Z * L;
    cudaError_t e = cudaMemsetAsync(dQacc, 0, (size_t)rows * 64 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    const int g = (int)((rows + 7) / 8 < 148L * 16 ? (rows + 7) / 8 : 148L * 16);
    flash_delta_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dO), reinterpret_cast<const __nv_bfloat16*>(O),
                                          ldo, B, nh, L, rows, delta);
    if ((r = pcm_launch_status())) return r;
    const long grid = (long)Z * ((S + 127) / 128);
    flash_bwd_kernel<<<(unsigned)grid, BWD_THREADS, BWD_SMEM, st>>>(tq, tk, tv, tdo, p);
    if ((r = pcm_launch_status())) return r;
    flash_dq_store_kernel<<<g, 256, 0, st>>>(dQacc, B, nh, L, rows, reinterpret_cast<__nv_bfloat16*>(dQ), ldq);
    return pcm_launch_status();
}
