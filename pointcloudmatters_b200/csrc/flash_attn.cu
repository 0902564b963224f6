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

#define MBWAIT(bar, parity) mbar_wait_backoff(bar, parity, 32)

constexpr uint32_t TILE_BYTES = 128 * 64 * 2;  // one [128 rows x 64] bf16 operand tile = 16 KB
constexpr float LOG2E = 1.4426950408889634f;

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
    long long* trace;  // debug: 64 clock64() stamps per CTA for the first trace_ctas CTAs (NULL = off)
    int trace_ctas;
};

#define TRACE(slot)                                                                                    \
    do {                                                                                               \
        if (p.trace != nullptr && (int)blockIdx.x < p.trace_ctas && (slot) < 64)                       \
            p.trace[(size_t)blockIdx.x * 64 + (slot)] = clock64();                                     \
    } while (0)

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

// Warp-private 4 KB staging tile (32 rows x 128 B, 16-byte pieces XOR-swizzled by row): a thread
// that owns one accumulator row parks it here, and the warp reads it back with lanes running
// along the row, so that every global store / reduction instruction touches 4 complete 128-byte
// lines instead of 32 different ones.
__device__ __forceinline__ void stage_put(uint8_t* stage, int lane, int piece, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4*>(stage + lane * 128 + ((piece ^ (lane & 7)) << 4)) = make_uint4(a, b, c, d);
}
__device__ __forceinline__ uint4 stage_get(const uint8_t* stage, int rl, int piece) {
    return *reinterpret_cast<const uint4*>(stage + rl * 128 + ((piece ^ (rl & 7)) << 4));
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
constexpr uint32_t FWD_SMEM_TILES = 5 * TILE_BYTES;   // Q, K, V, P (2 blocks)
constexpr uint32_t FWD_STAGE_BYTES = 4 * 4096;        // 4 softmax warps x 4 KB output staging
constexpr uint32_t FWD_SMEM = FWD_SMEM_TILES + FWD_STAGE_BYTES + 256 + 1024;

// Persistent: grid = min(#items, 2 x #SMs), two CTAs resident per SM; a CTA walks work items
// (batch*head z, 128-query tile) and all pipelines run across item boundaries: the next item's Q and
// first K / V tiles are loaded, and its first score tile is issued, while the softmax warps still
// normalise and store the current item's output.
template <bool DROPOUT>
__global__ void __launch_bounds__(FWD_THREADS, 2) flash_fwd_kernel(const __grid_constant__ CUtensorMap tq,
                                                                    const __grid_constant__ CUtensorMap tk,
                                                                    const __grid_constant__ CUtensorMap tv,
                                                                    const FlashParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = sm;
    uint8_t* sK = sm + TILE_BYTES;
    uint8_t* sV = sm + 2 * TILE_BYTES;
    uint8_t* sP = sm + 3 * TILE_BYTES;  // [128 x 128] bf16
    uint8_t* sStage = sm + FWD_SMEM_TILES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + FWD_SMEM_TILES + FWD_STAGE_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* k_empty = bars + 3;
    uint64_t* v_full = bars + 4;
    uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;
    uint64_t* p_full = bars + 7;
    uint64_t* pv_full = bars + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq_tiles = (p.L + 127) >> 7;
    const int n_kv = (p.S + 127) >> 7;
    const int n_items = p.B * p.nh * nq_tiles;

    if (threadIdx.x == 0) TRACE(0);
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv);
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        mbar_init(k_full, 1); mbar_init(k_empty, 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S = 0, TM_PV = 128;
    if (threadIdx.x == 0) TRACE(1);
    // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (warp-uniform control flow, one elected lane issues) =====
        uint32_t t = 0, n = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
            const int z = w / nq_tiles, q0 = (w - z * nq_tiles) << 7;
            MBWAIT(q_empty, (n & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(q_full, TILE_BYTES);
                tma_load_3d(sQ, &tq, q_full, 0, q0, z);
            }
            __syncwarp();
            for (int j = 0; j < n_kv; ++j, ++t) {
                MBWAIT(k_empty, (t & 1) ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(k_full, TILE_BYTES);
                    tma_load_3d(sK, &tk, k_full, 0, j << 7, z);
                }
                __syncwarp();
                MBWAIT(v_empty, (t & 1) ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(v_full, TILE_BYTES);
                    tma_load_3d(sV, &tv, v_full, 0, j << 7, z);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues =====
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
        const uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);
        // S = Q K_j^T of global tile tt (key tile j of the item whose Q is resident)
        auto issue_s = [&](uint32_t tt, int j) {
            const int nkv16 = (min(128, p.S - (j << 7)) + 15) & ~15;
            MBWAIT(k_full, tt & 1);
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, nkv16, false, false);
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + TM_S, make_smem_desc(aQ + k * 32, 16, 1024), make_smem_desc(aK + k * 32, 16, 1024), idesc,
                             k != 0 ? 1u : 0u);
                umma_commit(s_full);
                umma_commit(k_empty);
                if (j == n_kv - 1) umma_commit(q_empty);  // last score tile of the item: Q may be replaced
                TRACE(tt < 8 ? 8 + 2 * (int)tt : 64);
            }
            __syncwarp();
        };
        uint32_t t = 0, n = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
            if (n == 0) {
                MBWAIT(q_full, 0);
                issue_s(0, 0);
            }
            for (int j = 0; j < n_kv; ++j, ++t) {
                MBWAIT(p_full, t & 1);  // P_t in smem; S_t and PV_{t-1} have been read out of TMEM
                tc_fence_after();
                // PV_t goes to the tensor pipe BEFORE S_{t+1}: the pipe executes in order, so when the
                // softmax warps receive S_{t+1} the product they still have to collect (and the P tile
                // they are about to overwrite) is already finished with -- no second wait per tile
                MBWAIT(v_full, t & 1);
                tc_fence_after();
                const int ksteps = ((min(128, p.S - (j << 7)) + 15) & ~15) >> 4;
                if (elect_one_sync()) {
                    for (int ks = 0; ks < ksteps; ++ks)
                        umma_f16(tmem_base + TM_PV, make_smem_desc(aP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                                 make_smem_desc(aV + ks * 2048, 16384, 1024), idesc_pv, ks != 0 ? 1u : 0u);
                    umma_commit(pv_full);
                    umma_commit(v_empty);
                    TRACE(t < 8 ? 9 + 2 * (int)t : 64);
                }
                __syncwarp();
                if (j + 1 < n_kv) {
                    issue_s(t + 1, j + 1);
                } else if (w + (int)gridDim.x < n_items) {
                    MBWAIT(q_full, (n + 1) & 1);
                    issue_s(t + 1, 0);
                }
            }
        }
    } else {
        // ===== softmax warps: thread = query row =====
        // Straight-line, branch-free inner code (dropout is a template parameter; the key mask
        // costs a warp-uniform test per 32-column chunk) staged over whole chunks so that the 32
        // exp2 chains of a chunk are independent instruction streams.
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        const float sl2 = p.scale_log2;
        const uint32_t thr_hi = p.thr16 << 16;
        unsigned long long seed = 0;
        if (DROPOUT) seed = (p.seed_base ? *p.seed_base : 0ULL) * 0xD1342543DE82EF95ULL + p.seed_offset;
        uint8_t* stage = sStage + (warp - 2) * 4096;
        const bool tr = warp == 2 && lane == 0;
        uint32_t v[32];
        uint32_t t = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int z = w / nq_tiles, q0 = (w - z * nq_tiles) << 7;
            const int b = z / p.nh, h = z - b * p.nh;
            const int l = q0 + row;
            const bool warp_active = q0 + quad * 32 < p.L;  // warp-uniform
            // valid-key bits of the 32 keys starting at col0 (warp-uniform value)
            auto key_bits = [&](int col0) -> uint32_t {
                if (p.kpm == nullptr) {
                    const int r = p.S - col0;
                    return r >= 32 ? 0xFFFFFFFFu : (r <= 0 ? 0u : ((1u << r) - 1u));
                }
                const int col = col0 + lane;
                return __ballot_sync(PCM_FULL_MASK, col < p.S && p.kpm[(size_t)b * p.S + col] == 0);
            };
            uint32_t rseed = 0;
            if (DROPOUT) rseed = pcm_row_seed(seed, (unsigned long long)z * p.L + l);
            float o[64];
#pragma unroll
            for (int e = 0; e < 64; ++e) o[e] = 0.f;
            float m = -INFINITY;
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < n_kv; ++j, ++t) {
                const int nkv = min(128, p.S - (j << 7));
                MBWAIT(s_full, t & 1);
                tc_fence_after();
                if (tr) TRACE(t < 7 ? 24 + 5 * (int)t : 64);
                // Lazy reference maximum: only the first key tile of an item pays the max pass.  Later
                // tiles exponentiate against the reference m the row already has -- probabilities may
                // then exceed 1, which is exact as long as nothing overflows (O, the row sum and the
                // log-sum-exp all carry the same reference) -- and a warp falls back to the exact
                // max + rescale path only if a row sum leaves the safe range (scores that grow by more
                // than ~100 octaves between key tiles).
                auto row_max = [&]() -> float {  // raw (unscaled) maximum over the valid keys of this tile
                    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        if (c * 32 >= nkv) break;
                        tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                        const uint32_t bits = key_bits((j << 7) + c * 32);
                        tmem_ld_wait(v);
                        if (bits == 0xFFFFFFFFu) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(v[e]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                mx[e & 3] = fmaxf(mx[e & 3], ((bits >> e) & 1u) ? __uint_as_float(v[e]) : -INFINITY);
                        }
                    }
                    return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                };
                auto emit_p = [&](float m_ref) {  // P = exp2(scale*s - m_ref) -> row sums, dropout, bf16 tile in smem
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        if (c * 32 >= nkv) break;
                        tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                        const uint32_t bits = key_bits((j << 7) + c * 32);
                        uint32_t hb[4];
                        if (DROPOUT) {
                            const uint32_t grp0 = (uint32_t)((j << 7) + c * 32) >> 3;
#pragma unroll
                            for (int q = 0; q < 4; ++q) hb[q] = pcm_pair_bits(rseed, grp0 + q);
                        }
                        tmem_ld_wait(v);
                        float pr[32];
                        // exponent arguments and row sums on packed fp32 pairs (FFMA2 / FADD2: bit-identical, half the
                        // instructions); ex2 itself is one MUFU per element
                        const float2 sl22 = make_float2(sl2, sl2), nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            const float2 a = pcm_ffma2(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sl22, nm2);
                            pr[e] = fast_exp2(a.x);
                            pr[e + 1] = fast_exp2(a.y);
                        }
                        if (bits != 0xFFFFFFFFu) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) pr[e] = ((bits >> e) & 1u) ? pr[e] : 0.f;
                        }
                        {
                            float2 l01 = make_float2(ls[0], ls[1]), l23 = make_float2(ls[2], ls[3]);
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                l01 = pcm_fadd2(l01, make_float2(pr[e], pr[e + 1]));
                                l23 = pcm_fadd2(l23, make_float2(pr[e + 2], pr[e + 3]));
                            }
                            ls[0] = l01.x; ls[1] = l01.y; ls[2] = l23.x; ls[3] = l23.y;
                        }
                        if (DROPOUT) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint32_t x = hb[q];
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    pr[8 * q + k] = x >= thr_hi ? pr[8 * q + k] : 0.f;
                                    x = pcm_lcg_next(x);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            st_shared_v4(sP + sw128_off(c >> 1, row, (c & 1) * 4 + i), pack_bf16(pr[8 * i], pr[8 * i + 1]),
                                         pack_bf16(pr[8 * i + 2], pr[8 * i + 3]), pack_bf16(pr[8 * i + 4], pr[8 * i + 5]),
                                         pack_bf16(pr[8 * i + 6], pr[8 * i + 7]));
                    }
                };
                float m_ref = 0.f;
                bool row_has_ref = true;
                if (warp_active && j == 0) {  // exact: the scale is positive, max(scale * s) = scale * max(s)
                    const float m0 = row_max() * sl2;
                    row_has_ref = m0 != -INFINITY;
                    m_ref = row_has_ref ? m0 : 0.f;
                } else {
                    row_has_ref = m != -INFINITY;
                    m_ref = row_has_ref ? m : 0.f;
                }
                if (tr) TRACE(t < 7 ? 25 + 5 * (int)t : 64);
                // O += P_{t-1} V_{t-1}: with a common reference there is no rescale between tiles, so the
                // previous tile's product is collected AFTER this tile's probabilities are on their way
                // (its MMA was issued behind S_t and is still in flight when S_t arrives)
                bool pv_collected = (j == 0);
                auto collect_pv = [&]() {
                    if (!pv_collected) {
                        MBWAIT(pv_full, (t - 1) & 1);
                        tc_fence_after();
                        if (warp_active) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                tmem_ld_32x32b_x32(t_row + TM_PV + c * 32, v);
                                tmem_ld_wait(v);
#pragma unroll
                                for (int e = 0; e < 32; e += 2) {
                                    const float2 r = pcm_fadd2(make_float2(o[c * 32 + e], o[c * 32 + e + 1]),
                                                               make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
                                    o[c * 32 + e] = r.x; o[c * 32 + e + 1] = r.y;
                                }
                            }
                        }
                        pv_collected = true;
                    }
                };
                if (warp_active) {
                    const float lp0 = ls[0], lp1 = ls[1], lp2 = ls[2], lp3 = ls[3];
                    emit_p(m_ref);
                    const float tot = (ls[0] + ls[1]) + (ls[2] + ls[3]);
                    if (j > 0 && __any_sync(PCM_FULL_MASK, !(tot < 1.0e30f))) {
                        // exact path: true maximum, rescale what the row has accumulated, redo the tile
                        collect_pv();
                        const float m_new = fmaxf(row_has_ref ? m_ref : -INFINITY, row_max() * sl2);
                        const float m2 = m_new == -INFINITY ? 0.f : m_new;
                        const float alpha = row_has_ref ? fast_exp2(m_ref - m2) : 0.f;
#pragma unroll
                        for (int e = 0; e < 64; ++e) o[e] *= alpha;
                        ls[0] = lp0 * alpha; ls[1] = lp1 * alpha; ls[2] = lp2 * alpha; ls[3] = lp3 * alpha;
                        row_has_ref = m_new != -INFINITY;
                        m_ref = m2;
                        emit_p(m_ref);
                    }
                    // the reference only counts once a valid key has been seen
                    m = (row_has_ref || (ls[0] + ls[1]) + (ls[2] + ls[3]) > 0.f) ? m_ref : -INFINITY;
                }
                if (tr) TRACE(t < 7 ? 26 + 5 * (int)t : 64);
                collect_pv();
                if (tr) TRACE(t < 7 ? 27 + 5 * (int)t : 64);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                if (tr) TRACE(t < 7 ? 28 + 5 * (int)t : 64);
            }
            // ---- item epilogue: last PV tile, normalise, store (the MMA warp is already on the next item) ----
            MBWAIT(pv_full, (t - 1) & 1);
            tc_fence_after();
            if (warp_active) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld_32x32b_x32(t_row + TM_PV + c * 32, v);
                    tmem_ld_wait(v);
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const float2 r = pcm_fadd2(make_float2(o[c * 32 + e], o[c * 32 + e + 1]),
                                                   make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
                        o[c * 32 + e] = r.x; o[c * 32 + e + 1] = r.y;
                    }
                }
                const float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
                const float inv = lsum > 0.f ? p.keep_scale / lsum : 0.f;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    stage_put(stage, lane, i, pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                              pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rl = it * 4 + (lane >> 3), piece = lane & 7;
                    const int lr = q0 + quad * 32 + rl;
                    if (lr < p.L)
                        *reinterpret_cast<uint4*>(p.O + ((size_t)lr * p.B + b) * p.ldo + h * 64 + piece * 8) = stage_get(stage, rl, piece);
                }
                if (l < p.L) p.lse[(size_t)z * p.L + l] = lsum > 0.f ? m + log2f(lsum) : INFINITY;
            }
        }
        if (tr) TRACE(62);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
    if (threadIdx.x == 0) TRACE(63);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// delta[z, l] = sum_d dO[z, l, d] * O[l * B + b, h * 64 + d]; also zero-fills the fp32 dQ accumulator the
// backward kernel reduces into.  Four rows per warp: 8 lanes own one 64-element row (16-byte loads, a
// 3-step shuffle reduction), so a warp instruction moves 512 B instead of 128 B.
__global__ void __launch_bounds__(256) flash_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                                          int ldo, int B, int nh, int L, long rows, float* __restrict__ delta,
                                                          float* __restrict__ dq_acc) {
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r0 = wid * 4; r0 < rows; r0 += nwarps * 4) {
        const long r = r0 + sub;
        float s = 0.f;
        if (r < rows) {
            const int z = (int)(r / L), l = (int)(r - (long)z * L);
            const int b = z / nh, h = z - b * nh;
            const uint4 a = reinterpret_cast<const uint4*>(dO + (size_t)r * 64)[l8];
            const uint4 c = reinterpret_cast<const uint4*>(O + ((size_t)l * B + b) * ldo + h * 64)[l8];
            const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* cp = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 af = __bfloat1622float2(ap[k]), cf = __bfloat1622float2(cp[k]);
                s += af.x * cf.x + af.y * cf.y;
            }
            if (dq_acc != nullptr) {  // clears the dQ accumulator row (NULL: single key tile, dQ is stored directly)
                float4* q = reinterpret_cast<float4*>(dq_acc + (size_t)r * 64) + l8 * 2;
                q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        s += __shfl_xor_sync(PCM_FULL_MASK, s, 4);
        s += __shfl_xor_sync(PCM_FULL_MASK, s, 2);
        s += __shfl_xor_sync(PCM_FULL_MASK, s, 1);
        if (l8 == 0 && r < rows) delta[r] = s;
    }
}

// dQ token-major bf16 <- fp32 (Z, L, 64) accumulator          (four rows per warp, 32 B in / 16 B out per lane)
__global__ void __launch_bounds__(256) flash_dq_store_kernel(const float* __restrict__ acc, int B, int nh, int L, long rows,
                                                             __nv_bfloat16* __restrict__ dQ, int ldq) {
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = wid * 4 + sub; r < rows; r += nwarps * 4) {
        const int z = (int)(r / L), l = (int)(r - (long)z * L);
        const int b = z / nh, h = z - b * nh;
        const float4* a = reinterpret_cast<const float4*>(acc + (size_t)r * 64) + l8 * 2;
        const float4 a0 = a[0], a1 = a[1];
        uint4 pk;
        pk.x = pack_bf16(a0.x, a0.y); pk.y = pack_bf16(a0.z, a0.w);
        pk.z = pack_bf16(a1.x, a1.y); pk.w = pack_bf16(a1.z, a1.w);
        reinterpret_cast<uint4*>(dQ + ((size_t)l * B + b) * ldq + h * 64)[l8] = pk;
    }
}

constexpr int BWD_THREADS = 320;
constexpr uint32_t BWD_SMEM_TILES = 12 * TILE_BYTES;  // K[2], V[2], Q[2], dO[2], Pd (2 blocks), dS (2 blocks)
constexpr uint32_t BWD_STAGE_BYTES = 8 * 4096;        // 8 warps x 4 KB epilogue staging
constexpr uint32_t BWD_SMEM = BWD_SMEM_TILES + BWD_STAGE_BYTES + 256 + 1024;

// Persistent: grid = min(#items, #SMs); a CTA walks work items (batch*head z, 128-key tile jt) and,
// inside an item, the query tiles.  All pipelines run ACROSS item boundaries -- the producer
// prefetches the next item's K / V (2 stages) and Q / dO tiles while the current one computes, the
// MMA warp issues the next item's S / dP as soon as the last score tile of this one has been
// consumed, and the dK / dV / dQ epilogue of an item overlaps the tensor work of the next -- so
// TMEM allocation, barrier set-up and the first-load latency are paid once per CTA, not once per
// key tile (they were 25-50% of a CTA's life with one item per CTA).
template <bool DROPOUT>
__global__ void __launch_bounds__(BWD_THREADS, 1) flash_bwd_kernel(const __grid_constant__ CUtensorMap tq,
                                                                    const __grid_constant__ CUtensorMap tk,
                                                                    const __grid_constant__ CUtensorMap tv,
                                                                    const __grid_constant__ CUtensorMap tdo,
                                                                    const __grid_constant__ CUtensorMap tdq,
                                                                    const __grid_constant__ CUtensorMap tdk,
                                                                    const __grid_constant__ CUtensorMap tdv,
                                                                    const FlashParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sK = sm;                     // 2 stages
    uint8_t* sV = sm + 2 * TILE_BYTES;    // 2 stages
    uint8_t* sQ = sm + 4 * TILE_BYTES;    // 2 stages
    uint8_t* sdO = sm + 6 * TILE_BYTES;   // 2 stages
    uint8_t* sPd = sm + 8 * TILE_BYTES;   // [128 q x 128 kv] bf16
    uint8_t* sdS = sm + 10 * TILE_BYTES;
    uint8_t* sStage = sm + BWD_SMEM_TILES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + BWD_SMEM_TILES + BWD_STAGE_BYTES);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_empty = bars + 2;   // [2]
    uint64_t* qdo_full = bars + 4;   // [2]
    uint64_t* qdo_empty = bars + 6;  // [2]
    uint64_t* sdp_full = bars + 8;
    uint64_t* pds_full = bars + 9;
    uint64_t* dq_full = bars + 10;
    uint64_t* dkv_empty = bars + 11;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kv = (p.S + 127) >> 7;
    const int nq_tiles = (p.L + 127) >> 7;
    const int n_items = p.B * p.nh * n_kv;

    if (threadIdx.x == 0) TRACE(0);
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv); prefetch_tmap(&tdo);
        prefetch_tmap(&tdq); prefetch_tmap(&tdk); prefetch_tmap(&tdv);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
            mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1);
        }
        mbar_init(sdp_full, 1); mbar_init(pds_full, 8); mbar_init(dq_full, 1); mbar_init(dkv_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;
    if (threadIdx.x == 0) TRACE(1);
    // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (warp-uniform control flow, one elected lane issues) =====
        uint32_t t = 0, n = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
            const int z = w / n_kv, kv0 = (w - z * n_kv) << 7;
            const uint32_t ks = n & 1;
            MBWAIT(&kv_empty[ks], ((n >> 1) & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&kv_full[ks], 2 * TILE_BYTES);
                tma_load_3d(sK + ks * TILE_BYTES, &tk, &kv_full[ks], 0, kv0, z);
                tma_load_3d(sV + ks * TILE_BYTES, &tv, &kv_full[ks], 0, kv0, z);
            }
            __syncwarp();
            for (int i = 0; i < nq_tiles; ++i, ++t) {
                const uint32_t st = t & 1;
                MBWAIT(&qdo_empty[st], ((t >> 1) & 1) ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&qdo_full[st], 2 * TILE_BYTES);
                    tma_load_3d(sQ + st * TILE_BYTES, &tq, &qdo_full[st], 0, i << 7, z);
                    tma_load_3d(sdO + st * TILE_BYTES, &tdo, &qdo_full[st], 0, i << 7, z);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform control flow, one elected lane issues =====
        const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO);
        const uint32_t aPd = smem_u32(sPd), adS = smem_u32(sdS);
        const uint32_t idesc_t = make_idesc_bf16(128, 64, true, true);    // dV, dK: A and B MN-major
        const uint32_t idesc_dq = make_idesc_bf16(128, 64, false, true);  // dQ: A K-major, B MN-major
        // S = Q K^T and dP = dO V^T of global tile tt, which belongs to the item with key stage ks / width nkv16
        auto issue_sdp = [&](uint32_t tt, uint32_t ks, int nkv16) {
            const uint32_t st = tt & 1;
            MBWAIT(&qdo_full[st], (tt >> 1) & 1);
            tc_fence_after();
            const uint32_t idesc_s = make_idesc_bf16(128, nkv16, false, false);
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + TM_S, make_smem_desc(aQ + st * TILE_BYTES + k * 32, 16, 1024),
                             make_smem_desc(aK + ks * TILE_BYTES + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + TM_DP, make_smem_desc(adO + st * TILE_BYTES + k * 32, 16, 1024),
                             make_smem_desc(aV + ks * TILE_BYTES + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(sdp_full);
                TRACE(tt < 8 ? 8 + 2 * (int)tt : 64);
            }
            __syncwarp();
        };
        uint32_t t = 0, n = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++n) {
            const int z = w / n_kv, kv0 = (w - z * n_kv) << 7;
            const int nkv16 = (min(128, p.S - kv0) + 15) & ~15;
            const uint32_t ks = n & 1;
            if (n == 0) {
                MBWAIT(&kv_full[0], 0);
                issue_sdp(0, 0, nkv16);
            }
            for (int i = 0; i < nq_tiles; ++i, ++t) {
                const uint32_t st = t & 1;
                MBWAIT(pds_full, t & 1);  // Pd_t, dS_t in smem; S_t, dP_t, dQ_{t-1} read out of TMEM
                tc_fence_after();
                // look ahead: scores of the next tile -- of this item or of the first tile of the next one.
                // Preferred order is A(t+1) then B(t) (the softmax warps get their next tile sooner); if
                // the next Q / dO tile has not landed yet, B(t) goes first instead of idling behind the load.
                bool has_next = false;
                uint32_t ks2 = ks;
                int nkv16_2 = nkv16;
                if (i + 1 < nq_tiles) {
                    has_next = true;
                } else if (w + (int)gridDim.x < n_items) {
                    has_next = true;
                    const int w2 = w + (int)gridDim.x;
                    const int kv2 = (w2 - (w2 / n_kv) * n_kv) << 7;
                    ks2 = ks ^ 1;
                    nkv16_2 = (min(128, p.S - kv2) + 15) & ~15;
                    MBWAIT(&kv_full[ks2], ((n + 1) >> 1) & 1);
                }
                bool next_ready = false;
                if (has_next) {
                    uint32_t ok = 0;
                    if (lane == 0) ok = mbar_try_wait(&qdo_full[(t + 1) & 1], ((t + 1) >> 1) & 1) ? 1u : 0u;
                    next_ready = __shfl_sync(PCM_FULL_MASK, ok, 0) != 0;
                    if (next_ready) issue_sdp(t + 1, ks2, nkv16_2);
                }
                if (i == 0 && n > 0) {  // the previous item's dV / dK accumulators have been drained
                    MBWAIT(dkv_empty, (n - 1) & 1);
                    tc_fence_after();
                }
                const int qsteps = ((min(128, p.L - (i << 7)) + 15) & ~15) >> 4;  // query rows are the K dimension
                const int ksteps = nkv16 >> 4;
                if (elect_one_sync()) {
                    for (int q = 0; q < qsteps; ++q)
                        umma_f16(tmem_base + TM_DV, make_smem_desc(aPd + q * 2048, 16384, 1024),
                                 make_smem_desc(adO + st * TILE_BYTES + q * 2048, 16384, 1024), idesc_t, (i | q) != 0 ? 1u : 0u);
                    for (int q = 0; q < qsteps; ++q)
                        umma_f16(tmem_base + TM_DK, make_smem_desc(adS + q * 2048, 16384, 1024),
                                 make_smem_desc(aQ + st * TILE_BYTES + q * 2048, 16384, 1024), idesc_t, (i | q) != 0 ? 1u : 0u);
                    for (int q = 0; q < ksteps; ++q)
                        umma_f16(tmem_base + TM_DQ, make_smem_desc(adS + (q >> 2) * 16384 + (q & 3) * 32, 16, 1024),
                                 make_smem_desc(aK + ks * TILE_BYTES + q * 2048, 16384, 1024), idesc_dq, q != 0 ? 1u : 0u);
                    umma_commit(dq_full);
                    umma_commit(&qdo_empty[st]);
                    if (i == nq_tiles - 1) umma_commit(&kv_empty[ks]);
                    TRACE(t < 8 ? 9 + 2 * (int)t : 64);
                }
                __syncwarp();
                if (has_next && !next_ready) issue_sdp(t + 1, ks2, nkv16_2);
            }
        }
    } else {
        // ===== softmax / gradient warps: thread = (query row, 64-key half) =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        const float sl2 = p.scale_log2, sc = p.scale;
        const uint32_t thr_hi = p.thr16 << 16;
        const float keep_scale = p.keep_scale;
        unsigned long long seed = 0;
        if (DROPOUT) seed = (p.seed_base ? *p.seed_base : 0ULL) * 0xD1342543DE82EF95ULL + p.seed_offset;
        uint32_t v[32], w[32];
        const bool tr = warp == 2 && lane == 0;
        // The TMEM results of tile t-1 (its dQ tile and, if it closed an item, that item's dV / dK) are
        // drained while tile t's Pd / dS are already on their way to the MMA warp: the arithmetic of a
        // tile never waits for the previous tile's (or item's) gradient MMAs.
        int pz = 0, pkv0 = 0, pi = 0;  // previous tile: batch*head, key offset, query tile
        bool have_prev = false, prev_last = false;
        uint32_t dvp[16];
        auto drain_load = [&]() {  // TMEM -> registers (must complete before the barrier arrive)
            tmem_ld_32x32b_x32(t_row + TM_DQ + half * 32, v);
            tmem_ld_wait(v);
            if (prev_last) {
                tmem_ld_32x32b_x32(t_row + TM_DV + half * 32, w);
                tmem_ld_wait(w);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    dvp[q] = pack_bf16(__uint_as_float(w[2 * q]) * keep_scale, __uint_as_float(w[2 * q + 1]) * keep_scale);
                tmem_ld_32x32b_x32(t_row + TM_DK + half * 32, w);
                tmem_ld_wait(w);
            }
        };
        // registers -> SWIZZLE_128B staging tiles -> global through the TMA: the 128 x 64 fp32 dQ tile
        // leaves as two bulk tensor REDUCTIONS (add), the bf16 dV / dK tiles of a finished item as two
        // bulk tensor stores -- no red / st.global instruction streams in the softmax warps, rows past
        // L / S are clipped by the tensor maps.  Staging is shared by the 8 warps (named barrier 1).
        const bool issuer = (warp == 2 && lane == 0);
        const bool dq_direct = n_kv == 1;  // the launcher then passes the bf16 token-major dQ map in `tdq`
        auto drain_store = [&]() {
            if (issuer) bulk_wait_read_all();  // earlier bulk operations have finished reading the staging tiles
            named_bar_sync(1, 256);
            if (dq_direct) {
                // a single key tile per batch*head: this dQ tile is final -- bf16, token-major, one bulk tensor store
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    st_shared_v4(sStage + sw128_off(0, row, half * 4 + q),
                                 pack_bf16(__uint_as_float(v[8 * q]) * sc, __uint_as_float(v[8 * q + 1]) * sc),
                                 pack_bf16(__uint_as_float(v[8 * q + 2]) * sc, __uint_as_float(v[8 * q + 3]) * sc),
                                 pack_bf16(__uint_as_float(v[8 * q + 4]) * sc, __uint_as_float(v[8 * q + 5]) * sc),
                                 pack_bf16(__uint_as_float(v[8 * q + 6]) * sc, __uint_as_float(v[8 * q + 7]) * sc));
                fence_proxy_async();
                named_bar_sync(1, 256);
                if (issuer) {
                    const int pb = pz / p.nh, ph = pz - pb * p.nh;
                    tma_store_4d(&tdq, sStage, 0, ph, pb, pi << 7);
                    bulk_commit();
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    st_shared_v4(sStage + sw128_off(half, row, c), __float_as_uint(__uint_as_float(v[4 * c]) * sc),
                                 __float_as_uint(__uint_as_float(v[4 * c + 1]) * sc), __float_as_uint(__uint_as_float(v[4 * c + 2]) * sc),
                                 __float_as_uint(__uint_as_float(v[4 * c + 3]) * sc));
                fence_proxy_async();
                named_bar_sync(1, 256);
                if (issuer) {
                    tma_reduce_add_3d(&tdq, sStage, 0, pi << 7, pz);
                    tma_reduce_add_3d(&tdq, sStage + 16384, 32, pi << 7, pz);
                    bulk_commit();
                }
            }
            if (prev_last) {
                if (issuer) bulk_wait_read_all();
                named_bar_sync(1, 256);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    st_shared_v4(sStage + sw128_off(0, row, half * 4 + q), dvp[4 * q], dvp[4 * q + 1], dvp[4 * q + 2], dvp[4 * q + 3]);
                    st_shared_v4(sStage + sw128_off(1, row, half * 4 + q),
                                 pack_bf16(__uint_as_float(w[8 * q]) * sc, __uint_as_float(w[8 * q + 1]) * sc),
                                 pack_bf16(__uint_as_float(w[8 * q + 2]) * sc, __uint_as_float(w[8 * q + 3]) * sc),
                                 pack_bf16(__uint_as_float(w[8 * q + 4]) * sc, __uint_as_float(w[8 * q + 5]) * sc),
                                 pack_bf16(__uint_as_float(w[8 * q + 6]) * sc, __uint_as_float(w[8 * q + 7]) * sc));
                }
                fence_proxy_async();
                named_bar_sync(1, 256);
                if (issuer) {
                    const int pb = pz / p.nh, ph = pz - pb * p.nh;
                    tma_store_4d(&tdv, sStage, 0, ph, pb, pkv0);
                    tma_store_4d(&tdk, sStage + 16384, 0, ph, pb, pkv0);
                    bulk_commit();
                }
            }
        };

        uint32_t t = 0, n = 0;
        for (int wi = blockIdx.x; wi < n_items; wi += gridDim.x, ++n) {
            const int z = wi / n_kv, kv0 = (wi - z * n_kv) << 7;
            const int b = z / p.nh;
            const int nkv16 = (min(128, p.S - kv0) + 15) & ~15;
            // valid-key bits of this thread's two 32-key chunks
            uint32_t kb[2];
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                const int col = kv0 + (half * 2 + c2) * 32 + lane;
                const bool valid = col < p.S && !(p.kpm != nullptr && p.kpm[(size_t)b * p.S + col] != 0);
                kb[c2] = __ballot_sync(PCM_FULL_MASK, valid);
            }
            // per-row statistics of the NEXT tile are fetched one iteration ahead
            float lse_n = row < p.L ? p.lse[(size_t)z * p.L + row] : INFINITY;
            float delta_n = row < p.L ? p.delta[(size_t)z * p.L + row] : 0.f;
            for (int i = 0; i < nq_tiles; ++i, ++t) {
                const int nq = min(128, p.L - (i << 7));
                const int nq16 = (nq + 15) & ~15;
                const bool warp_active = quad * 32 < nq16;  // warp-uniform: rows this warp owns are read by the MMAs
                const int l = (i << 7) + row;
                const bool row_valid = row < nq;
                const float lse_r = lse_n, delta_r = delta_n;
                if (i + 1 < nq_tiles) {
                    const int l2 = l + 128;
                    lse_n = l2 < p.L ? p.lse[(size_t)z * p.L + l2] : INFINITY;
                    delta_n = l2 < p.L ? p.delta[(size_t)z * p.L + l2] : 0.f;
                }
                uint32_t rseed = 0;
                if (DROPOUT) rseed = pcm_row_seed(seed, (unsigned long long)z * p.L + l);
                // Pd / dS smem (and the TMEM results of tile t-1) become free when tile t-1's MMAs complete;
                // that is waited for only in front of the first shared-memory store, i.e. after half of the
                // tile's arithmetic
                bool prev_done = !have_prev;
                auto wait_prev = [&]() {
                    if (!prev_done) {
                        MBWAIT(dq_full, (t - 1) & 1);
                        tc_fence_after();
                        prev_done = true;
                    }
                };
                MBWAIT(sdp_full, t & 1);
                tc_fence_after();
                if (tr) TRACE(t < 7 ? 24 + 5 * (int)t : 64);
                if (warp_active) {
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int c = half * 2 + c2;
                        if (c * 32 < nkv16) {
                            tmem_ld_32x32b_x32(t_row + TM_S + c * 32, v);
                            tmem_ld_32x32b_x32(t_row + TM_DP + c * 32, w);
                            const uint32_t bits = row_valid ? kb[c2] : 0u;
                            uint32_t hb[4];
                            if (DROPOUT) {
                                const uint32_t grp0 = (uint32_t)(kv0 + c * 32) >> 3;
#pragma unroll
                                for (int q = 0; q < 4; ++q) hb[q] = pcm_pair_bits(rseed, grp0 + q);
                            }
                            tmem_ld_wait(v);
                            tmem_ld_wait(w);
                            float pr[32], g[32];
                            uint32_t pd_pk[16], ds_pk[16];
                            // exponent arguments and the dS arithmetic below on packed fp32 pairs (FFMA2 / FMUL2: bit-identical)
                            const float2 sl22 = make_float2(sl2, sl2), nl2 = make_float2(-lse_r, -lse_r);
#pragma unroll
                            for (int e = 0; e < 32; e += 2) {
                                const float2 a = pcm_ffma2(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sl22, nl2);
                                pr[e] = fast_exp2(a.x);
                                pr[e + 1] = fast_exp2(a.y);
                                g[e] = __uint_as_float(w[e]);
                                g[e + 1] = __uint_as_float(w[e + 1]);
                            }
                            if (bits != 0xFFFFFFFFu) {  // masked keys, rows past L, columns past nkv16 (uninitialised TMEM)
#pragma unroll
                                for (int e = 0; e < 32; ++e) {
                                    pr[e] = ((bits >> e) & 1u) ? pr[e] : 0.f;
                                    g[e] = ((bits >> e) & 1u) ? g[e] : 0.f;
                                }
                            }
                            // dropout: Pd = keep ? P : 0 and dS = P * (keep ? keep_scale * dP : 0 - delta); the
                            // keep_scale of Pd and the softmax scale of dS are folded into the epilogues
                            if (DROPOUT) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint32_t x = hb[q];
#pragma unroll
                                    for (int k = 0; k < 8; ++k) {
                                        const bool keep = x >= thr_hi;
                                        x = pcm_lcg_next(x);
                                        g[8 * q + k] = keep ? g[8 * q + k] : 0.f;
                                        w[8 * q + k] = __float_as_uint(keep ? pr[8 * q + k] : 0.f);
                                    }
                                }
#pragma unroll
                                for (int q = 0; q < 16; ++q)
                                    pd_pk[q] = pack_bf16(__uint_as_float(w[2 * q]), __uint_as_float(w[2 * q + 1]));
                                const float2 ks2 = make_float2(keep_scale, keep_scale), nd2 = make_float2(-delta_r, -delta_r);
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    const float2 d = pcm_fmul2(make_float2(pr[2 * q], pr[2 * q + 1]),
                                                               pcm_ffma2(make_float2(g[2 * q], g[2 * q + 1]), ks2, nd2));
                                    ds_pk[q] = pack_bf16(d.x, d.y);
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < 16; ++q) pd_pk[q] = pack_bf16(pr[2 * q], pr[2 * q + 1]);
                                const float2 nd2 = make_float2(-delta_r, -delta_r);
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    const float2 d = pcm_fmul2(make_float2(pr[2 * q], pr[2 * q + 1]),
                                                               pcm_fadd2(make_float2(g[2 * q], g[2 * q + 1]), nd2));
                                    ds_pk[q] = pack_bf16(d.x, d.y);
                                }
                            }
                            if (tr && c2 == 0) TRACE(t < 7 ? 25 + 5 * (int)t : 64);
                            wait_prev();
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t off = sw128_off(c >> 1, row, (c & 1) * 4 + q);
                                st_shared_v4(sPd + off, pd_pk[4 * q], pd_pk[4 * q + 1], pd_pk[4 * q + 2], pd_pk[4 * q + 3]);
                                st_shared_v4(sdS + off, ds_pk[4 * q], ds_pk[4 * q + 1], ds_pk[4 * q + 2], ds_pk[4 * q + 3]);
                            }
                        }
                    }
                }
                wait_prev();
                if (tr) TRACE(t < 7 ? 26 + 5 * (int)t : 64);
                fence_proxy_async();
                if (have_prev) drain_load();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(pds_full);
                    if (have_prev && prev_last) mbar_arrive(dkv_empty);  // the closed item's accumulators are drained
                }
                if (tr) TRACE(t < 7 ? 27 + 5 * (int)t : 64);
                if (have_prev) drain_store();
                if (tr) TRACE(t < 7 ? 28 + 5 * (int)t : 64);
                pz = z; pkv0 = kv0; pi = i;
                prev_last = (i == nq_tiles - 1);
                have_prev = true;
            }
        }
        if (have_prev) {  // the very last tile of this CTA
            MBWAIT(dq_full, (t - 1) & 1);
            tc_fence_after();
            drain_load();
            drain_store();
        }
        if (issuer) bulk_wait_all();  // bulk reductions / stores complete before the CTA retires
        if (tr) TRACE(62);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) TRACE(63);
}

long long* g_trace_buf = nullptr;
int g_trace_ctas = 0;

int head_split_map(const void* ptr, int Z, int rows, CUtensorMap* out) {
    const uint64_t dims[3] = {64, (uint64_t)rows, (uint64_t)Z};
    const uint64_t strides[2] = {128, (uint64_t)rows * 128};
    const uint32_t box[3] = {64, 128, 1};
    return tensor_map_bf16(ptr, 3, dims, strides, box, out);
}

int fill_common(FlashParams& p, int B, int nh, int L, int S, const unsigned char* kpm, float scale, float p_drop,
                const unsigned long long* seed_base, unsigned long long seed_offset) {
    if (B <= 0 || nh <= 0 || L <= 0 || S <= 0 || !(scale > 0.f)) return PCM_EINVAL;
    if (p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    p.B = B; p.nh = nh; p.L = L; p.S = S; p.kpm = kpm;
    p.scale = scale; p.scale_log2 = scale * LOG2E;
    p.thr16 = p_drop > 0.f ? pcm_drop_thr16(p_drop) : 0u;
    p.keep_scale = p.thr16 ? pcm_keep_scale(p.thr16) : 1.0f;
    p.seed_base = seed_base; p.seed_offset = seed_offset;
    p.trace = g_trace_buf; p.trace_ctas = g_trace_ctas;
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
        cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(flash_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const long items = (long)Z * ((L + 127) / 128);
    const long grid = items < 2L * num_sms ? items : 2L * num_sms;
    cudaError_t le = p.thr16 ? pcm_launch(flash_fwd_kernel<true>, dim3((unsigned)grid), dim3(FWD_THREADS), FWD_SMEM, pcm_cu_stream(stream), tq, tk, tv, p)
                             : pcm_launch(flash_fwd_kernel<false>, dim3((unsigned)grid), dim3(FWD_THREADS), FWD_SMEM, pcm_cu_stream(stream), tq, tk, tv, p);
    if (le != cudaSuccess) return (int)le;
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
    CUtensorMap tdq, tdk, tdv;
    // One key tile per batch*head (S <= 128: decoder self-attention, CVAE encoder): every dQ tile is final when its
    // query tile has been processed, so the kernel stores it directly (bf16, token-major) -- no zero-fill of the fp32
    // accumulator, no reduction, no conversion kernel.
    const bool dq_direct = S <= 128 && !(ldq % 8) && !(reinterpret_cast<uintptr_t>(dQ) & 15);
    if (dq_direct) {  // element (d, h, b, l) at ((l * B + b) * ldq + h * 64 + d)
        const uint64_t dims[4] = {64, (uint64_t)nh, (uint64_t)B, (uint64_t)L};
        const uint64_t strides[3] = {128, (uint64_t)ldq * 2, (uint64_t)B * ldq * 2};
        const uint32_t box[4] = {64, 1, 1, 128};
        if ((r = tensor_map(dQ, false, 4, dims, strides, box, &tdq))) return r;
    } else {  // fp32 dQ accumulator (Z, L, 64): boxes of 32 columns x 128 rows
        const uint64_t dims[3] = {64, (uint64_t)L, (uint64_t)Z};
        const uint64_t strides[2] = {256, (uint64_t)L * 256};
        const uint32_t box[3] = {32, 128, 1};
        if ((r = tensor_map(dQacc, true, 3, dims, strides, box, &tdq))) return r;
    }
    {   // token-major bf16 dK / dV: element (d, h, b, s) at ((s * B + b) * ldkv + h * 64 + d)
        const uint64_t dims[4] = {64, (uint64_t)nh, (uint64_t)B, (uint64_t)S};
        const uint64_t strides[3] = {128, (uint64_t)ldkv * 2, (uint64_t)B * ldkv * 2};
        const uint32_t box[4] = {64, 1, 1, 128};
        if ((r = tensor_map(dK, false, 4, dims, strides, box, &tdk))) return r;
        if ((r = tensor_map(dV, false, 4, dims, strides, box, &tdv))) return r;
    }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(flash_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    cudaStream_t st = pcm_cu_stream(stream);
    const long rows = (long)Z * L;
    if ((ldo % 8) || (ldq % 8) || (reinterpret_cast<uintptr_t>(O) & 15) || (reinterpret_cast<uintptr_t>(dQ) & 15) ||
        (reinterpret_cast<uintptr_t>(dO) & 15))
        return PCM_EUNSUPPORTED;  // 16-byte row accesses of the pre / post kernels
    const long row_ctas = (rows + 31) / 32;  // 8 warps x 4 rows per CTA
    const int g = (int)(row_ctas < 148L * 8 ? (row_ctas > 0 ? row_ctas : 1) : 148L * 8);
    cudaError_t le = pcm_launch(flash_delta_kernel, dim3(g), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(dO),
                                reinterpret_cast<const __nv_bfloat16*>(O), ldo, B, nh, L, rows, delta, dq_direct ? nullptr : dQacc);
    if (le != cudaSuccess) return (int)le;
    if ((r = pcm_launch_status())) return r;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const long items = (long)Z * ((S + 127) / 128);
    const long grid = items < num_sms ? items : num_sms;
    le = p.thr16 ? pcm_launch(flash_bwd_kernel<true>, dim3((unsigned)grid), dim3(BWD_THREADS), BWD_SMEM, st, tq, tk, tv, tdo, tdq, tdk, tdv, p)
                 : pcm_launch(flash_bwd_kernel<false>, dim3((unsigned)grid), dim3(BWD_THREADS), BWD_SMEM, st, tq, tk, tv, tdo, tdq, tdk, tdv, p);
    if (le != cudaSuccess) return (int)le;
    if ((r = pcm_launch_status())) return r;
    if (dq_direct) return PCM_OK;
    le = pcm_launch(flash_dq_store_kernel, dim3(g), dim3(256), 0, st, (const float*)dQacc, B, nh, L, rows,
                    reinterpret_cast<__nv_bfloat16*>(dQ), ldq);
    if (le != cudaSuccess) return (int)le;
    return pcm_launch_status();
}

// Debug aid (tools/flash_trace.py): the first n_ctas CTAs of subsequent attention launches write 64
// clock64() stamps each (phase boundaries of the producer / MMA / softmax roles) to `buf`
// (device memory, n_ctas * 64 int64).  buf = NULL switches tracing off.
PCM_API int pcm_flash_attn_debug_trace(long long* buf, int n_ctas) {
    g_trace_buf = buf;
    g_trace_ctas = buf ? n_ctas : 0;
    return PCM_OK;
}
