// gemm_tcgen05.cu -- bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (sm_100a).
//
//     C[m, n] (+)= sum_k A(m, k) * B(n, k)   (+ bias[n]) (ReLU)        C: fp32 or bf16
//
// Hand-written tcgen05 / TMEM / TMA kernel (inline PTX, no CUTLASS):
//   * warp 0 (one elected lane): TMA producer -- cp.async.bulk.tensor.2d tiles into a 128B-swizzled
//     shared-memory ring, completion counted on mbarriers (expect_tx);
//   * warp 1 (one elected lane): issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BLOCK_N x 16)
//     with the fp32 accumulator in TENSOR MEMORY, releases ring slots with tcgen05.commit;
//     the same warp owns tcgen05.alloc / dealloc;
//   * warps 2-9: epilogue (two warps per TMEM lane quadrant, each draining half of the tile's
//     columns: short-K GEMMs are epilogue-latency-bound) -- tcgen05.ld (32 lanes x 32 columns per warp) -> bias / ReLU / cast ->
//     swizzled shared-memory strip -> fully coalesced 128-byte row-segment stores (or coalesced
//     red.global.add.f32 for split-K / gradient accumulation).
// Both operands can be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the
// second form lets the backward GEMMs (dX = dY W, dW = dY^T X) read activations / weights in the
// layout they already have, with no transpose pass.  Shared-memory layouts are the canonical UMMA
// SWIZZLE_128B layouts (K-major: 8-row x 128 B atoms, SBO = 1024 B; MN-major: 64-element x 8-row
// atoms, SBO = 1024 B, LBO = BLOCK_K * 128 B) which are exactly what a SWIZZLE_128B TMA box writes.
// Persistent: one CTA per SM walks output tiles; a 6-stage ring (192 KB) keeps TMA ahead of the
// MMA warp across tile boundaries and two TMEM accumulators overlap a tile's epilogue with the
// next tile's main loop.
#include "common.cuh"
#include "tcgen05_ptx.cuh"

#include <cuda.h>
#include <mutex>
#include <unordered_map>

namespace {

using namespace pcm_tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
template <int BLOCK_N> struct StagesFor { static constexpr int value = BLOCK_N == 256 ? 4 : 6; };  // <= 192 KB ring
constexpr int NUM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quadrant)
constexpr int NUM_EPI_WARPS = 8;

struct GemmParams {
    int M, N, K;             // per-batch problem size
    int k_blocks_per_split;  // split-K: slice ks covers [ks * kbps, min((ks+1) * kbps, kblocks))
    int split_k, batch;      // tile index enumerates (batch, k-slice, m-block, n-block)
    // batching: operand row offset per batch, in rows of the operand's 2-D tensor (for an
    // MN-major operand the rows are the K dimension)
    long long a_batch_rows, b_batch_rows;
    void* C;
    int ldc;
    // output addressing: 0 plain (row = z * c_batch_rows + m), 1 head-split (token-major rows
    // r = l * hs_B + b, column = h * 64 + d  ->  ((b * hs_nh + h) * hs_L + l) * 64 + d),
    // 2 head-merge (batch z = b * hs_nh + h, row l, column d -> (l * hs_B + b) * ldc + h * 64 + d)
    int c_mode;
    long long c_batch_rows;
    int hs_B, hs_nh, hs_L;
    const float* bias;
    int relu;
    int c_bf16;
    int atomic;  // red.global.add.f32 (split-K / gradient accumulation); C must be fp32
    float alpha;  // C = alpha * acc (+ bias)
    // second A operand: output columns n >= a2_from_col (a multiple of BLOCK_N) are computed from tensor map A2 instead
    // of A -- the fused Q|K|V in-projection of nn.MultiheadAttention reads bf16(x + pos) for Q, K and bf16(x) for V in
    // ONE launch.  INT_MAX = off.
    int a2_from_col;
    // head-split output in parts (c_mode 1): column n belongs to part n / hs_part_cols, whose (B, nh, L, 64) tensor
    // starts hs_part_stride elements after the previous part's (Q | K | V buffers).  0 = a single part.
    int hs_part_cols;
    long long hs_part_stride;
};

// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// Tile order: n fastest, then m, then (batch, k-slice) -- CTAs running at the same time share the
// A row-block and the whole (small) B operand through L2.  Two TMEM accumulators (2 x BLOCK_N
// columns) let the epilogue warps drain tile i while the MMA warp already accumulates tile i+1.
enum { EPI_F32 = 0, EPI_BF16 = 1, EPI_ATOMIC = 2 };

// head-split destination offset of output column `col` relative to the row's offset: (part, head, d)
__device__ __forceinline__ size_t hs_col_off(const GemmParams& p, int col) {
    size_t off = 0;
    if (p.hs_part_cols > 0) {
        const int part = col / p.hs_part_cols;
        col -= part * p.hs_part_cols;
        off = (size_t)part * (size_t)p.hs_part_stride;
    }
    return off + (size_t)(col >> 6) * p.hs_L * 64 + (col & 63);
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                      const __grid_constant__ CUtensorMap tmap_b,
                                                                      const __grid_constant__ CUtensorMap tmap_a2,
                                                                      const GemmParams p) {
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int STAGES = StagesFor<BLOCK_N>::value;
    constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // 128, 256 or 512: power of two >= 32
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=F32, A=B=BF16
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint8_t* epi_stage = tiles + STAGES * STAGE_BYTES + 256;  // 8 warps x 4 KB, 128-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles_mn = ((p.M + BLOCK_M - 1) / BLOCK_M) * tiles_n;
    const int total_tiles = tiles_mn * p.split_k * p.batch;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        if (p.a2_from_col < p.N) prefetch_tmap(&tmap_a2);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above touched only parameters, shared and tensor memory: let the next kernel start
    // its own prologue, then wait for the producers of our operands (programmatic dependent launch)
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (warp-uniform control flow, one elected lane issues) =====
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int zs = t / tiles_mn, r = t - zs * tiles_mn;
            const int m_blk = r / tiles_n, n_blk = r - m_blk * tiles_n;
            const int zb = zs / p.split_k, ks = zs - zb * p.split_k;
            const int a_off = (int)(zb * p.a_batch_rows), b_off = (int)(zb * p.b_batch_rows);
            const int kb0 = ks * p.k_blocks_per_split;
            const int kb1 = min(kblocks_total, kb0 + p.k_blocks_per_split);
            const CUtensorMap* amap = n_blk * BLOCK_N >= p.a2_from_col ? &tmap_a2 : &tmap_a;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (elect_one_sync()) {
                    uint8_t* sa = tiles + s * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    const int k0 = kb * BLOCK_K;
                    if (!A_MN) {
                        tma_load_2d(sa, amap, &full_bar[s], k0, a_off + m_blk * BLOCK_M);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_M / 64; ++c)
                            tma_load_2d(sa + c * (BLOCK_K * 128), amap, &full_bar[s], m_blk * BLOCK_M + c * 64, a_off + k0);
                    }
                    if (!B_MN) {
                        tma_load_2d(sb, &tmap_b, &full_bar[s], k0, b_off + n_blk * BLOCK_N);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_N / 64; ++c)
                            tma_load_2d(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[s], n_blk * BLOCK_N + c * 64, b_off + k0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues.
        // With an `if (lane == 0)` region the compiler wraps every UTCHMMA in a per-thread waterfall
        // loop (R2UR + ELECT + BRA.U.ANY, ~20 instructions); warp-uniform descriptors stay in uniform
        // registers and the MMAs of a k-block issue back to back.
        uint32_t it = 0;
        int i = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
            const int zs = t / tiles_mn;
            const int ks = zs % p.split_k;
            const int kb0 = ks * p.k_blocks_per_split;
            const int nkb = min(kblocks_total, kb0 + p.k_blocks_per_split) - kb0;
            const int acc = i & 1;
            mbar_wait(&tmem_empty_bar[acc], (((uint32_t)i >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + s * STAGE_BYTES);
                const uint32_t sb = sa + A_BYTES;
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // K-major: +16 elements = +32 B inside the swizzle row; MN-major: +16 rows of 128 B
                        const uint64_t da = A_MN ? make_smem_desc(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                                 : make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t db = B_MN ? make_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                                 : make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
                        umma_f16(tmem_d, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
                    if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps (2..5): TMEM lane quadrant = warp % 4 =====
        // Each warp drains its own 32 accumulator rows in 32-column strips.  The TMEM loads are
        // software-pipelined (strip q+1 is in flight while strip q is processed: with one epilogue
        // warp per SM sub-partition nothing else hides the tcgen05.ld latency).  A strip goes
        // registers -> (alpha, bias, ReLU, cast) -> a private swizzled 4 KB shared-memory strip ->
        // read back transposed, so every global store instruction covers complete 128-byte row
        // segments instead of 32 different cache lines.
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
        const int chalf = (warp - 2) >> 2;         // which half of the tile's columns it drains
        uint8_t* stage = epi_stage + (warp - 2) * 4096;
        constexpr int NLD = BLOCK_N / 64;          // 32-column loads per warp
        constexpr int ESZ = EPI == EPI_BF16 ? 2 : 4;
        int i = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
            const int zs = t / tiles_mn, r = t - zs * tiles_mn;
            const int m_blk = r / tiles_n, n_blk = r - m_blk * tiles_n;
            const int zb = zs / p.split_k;
            const int acc = i & 1;
            const int row_base = m_blk * BLOCK_M + quad * 32;  // first row of this warp inside the batch's M
            const int cbase = chalf * (BLOCK_N / 2);  // first tile column of this warp
            const int col_lim = min(p.N, n_blk * BLOCK_N + cbase + BLOCK_N / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + cbase);
            // per-row destination offsets (elements) of this warp's 32 rows, lane rl holds row rl
            size_t row_dst;
            {
                const int row = row_base + lane;
                if (p.c_mode == 0) row_dst = (size_t)(zb * p.c_batch_rows + row) * p.ldc;
                else if (p.c_mode == 1) { const int l = row / p.hs_B, b = row - l * p.hs_B; row_dst = ((size_t)b * p.hs_nh * p.hs_L + l) * 64; }
                else { const int b = zb / p.hs_nh, h = zb - b * p.hs_nh; row_dst = ((size_t)row * p.hs_B + b) * p.ldc + h * 64; }
            }
            mbar_wait(&tmem_full_bar[acc], ((uint32_t)i >> 1) & 1);
            tc_fence_after();

            auto process = [&](uint32_t (&v)[32], int q) {
                const int col0 = n_blk * BLOCK_N + cbase + q * 32;
                if (col0 >= p.N) return;  // warp-uniform
                float f[32];
                if (p.bias != nullptr) {  // one coalesced load per strip, broadcast by shuffles
                    const float bl = col0 + lane < p.N ? __ldg(p.bias + col0 + lane) : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha + __shfl_sync(PCM_FULL_MASK, bl, j);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (EPI == EPI_BF16) {
                    // two consecutive loads (64 columns) share one 128-byte-per-row strip
                    const int half = q & 1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 pk;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(f[c * 8 + 0], f[c * 8 + 1]);
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]);
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&b0); pk.y = *reinterpret_cast<uint32_t*>(&b1);
                        pk.z = *reinterpret_cast<uint32_t*>(&b2); pk.w = *reinterpret_cast<uint32_t*>(&b3);
                        const int chunk = half * 4 + c;
                        *reinterpret_cast<uint4*>(stage + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = pk;
                    }
                    if (half == 0 && col0 + 32 < col_lim && q + 1 < NLD) return;  // wait for the second half of the strip
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                            make_float4(f[c * 4 + 0], f[c * 4 + 1], f[c * 4 + 2], f[c * 4 + 3]);
                }
                __syncwarp();
                const int scol0 = EPI == EPI_BF16 ? (n_blk * BLOCK_N + cbase + (q & ~1) * 32) : col0;  // first column of the strip
                if (EPI == EPI_ATOMIC) {
#pragma unroll 4
                    for (int rl = 0; rl < 32; ++rl) {
                        const size_t rd = __shfl_sync(PCM_FULL_MASK, row_dst, rl);
                        const float x = *reinterpret_cast<const float*>(stage + rl * 128 + (((lane >> 2) ^ (rl & 7)) << 4) + ((lane & 3) << 2));
                        if (row_base + rl < p.M && scol0 + lane < col_lim) {
                            const size_t dst = p.c_mode == 1 ? rd + hs_col_off(p, scol0 + lane) : rd + scol0 + lane;
                            atomicAdd(reinterpret_cast<float*>(p.C) + dst, x);
                        }
                    }
                } else {
#pragma unroll
                    for (int it8 = 0; it8 < 8; ++it8) {
                        const int rl = it8 * 4 + (lane >> 3);
                        const int chunk = lane & 7;
                        const size_t rd = __shfl_sync(PCM_FULL_MASK, row_dst, rl);
                        const uint4 pk = *reinterpret_cast<const uint4*>(stage + rl * 128 + ((chunk ^ (rl & 7)) << 4));
                        const int ecol = scol0 + chunk * (16 / ESZ);  // first element column of this 16-byte piece
                        if (row_base + rl < p.M && ecol < col_lim) {
                            const size_t dst = p.c_mode == 1 ? rd + hs_col_off(p, ecol) : rd + ecol;
                            uint8_t* g = reinterpret_cast<uint8_t*>(p.C) + dst * ESZ;
                            if (ecol + 16 / ESZ <= col_lim && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
                                *reinterpret_cast<uint4*>(g) = pk;
                            } else if (EPI == EPI_BF16) {
                                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&pk);
                                for (int qq = 0; qq < 8; ++qq)
                                    if (ecol + qq < col_lim) reinterpret_cast<__nv_bfloat16*>(g)[qq] = e[qq];
                            } else {
                                const float* e = reinterpret_cast<const float*>(&pk);
                                for (int qq = 0; qq < 4; ++qq)
                                    if (ecol + qq < col_lim) reinterpret_cast<float*>(g)[qq] = e[qq];
                            }
                        }
                    }
                }
                __syncwarp();
            };

            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(taddr, va);
#pragma unroll 1
            for (int q = 0; q < NLD; q += 2) {
                tmem_ld_wait(va);
                if (q + 1 < NLD) tmem_ld_32x32b_x32(taddr + (uint32_t)((q + 1) * 32), vb);
                process(va, q);
                if (q + 1 < NLD) {
                    tmem_ld_wait(vb);
                    if (q + 2 < NLD) tmem_ld_32x32b_x32(taddr + (uint32_t)((q + 2) * 32), va);
                    process(vb, q + 1);
                }
            }
            // this warp has read its accumulator quadrant: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    uint64_t inner, outer, ld;
    uint32_t box_inner, box_outer;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        auto mix = [&](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
        mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer);
        return h;
    }
};

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with row pitch `ld` elements,
// SWIZZLE_128B boxes of box_inner (= 64) x box_outer elements, zero fill out of bounds.
int get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer, CUtensorMap* out) {
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    static std::mutex mu;
    MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[2] = {inner, outer};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 700 + (int)r;
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return 0;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
int launch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const GemmParams& p, int split_k, int batch,
               cudaStream_t st) {
    constexpr size_t SMEM = StagesFor<BLOCK_N>::value * (BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2) + 1024 + 256 + NUM_EPI_WARPS * 4096;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, EPI>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
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
    const long tiles = (long)((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N) * split_k * batch;
    const int grid = (int)(tiles < num_sms ? tiles : num_sms);
    cudaError_t le = pcm_launch(gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, EPI>, dim3(grid), dim3(NUM_THREADS), SMEM, st, ta, tb, ta2, p);
    if (le != cudaSuccess) return (int)le;
    return pcm_launch_status();
}

template <int BLOCK_N, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const GemmParams& p, int split_k, int batch,
           cudaStream_t st) {
    if (p.atomic) return launch_epi<BLOCK_N, A_MN, B_MN, EPI_ATOMIC>(ta, tb, ta2, p, split_k, batch, st);
    if (p.c_bf16) return launch_epi<BLOCK_N, A_MN, B_MN, EPI_BF16>(ta, tb, ta2, p, split_k, batch, st);
    return launch_epi<BLOCK_N, A_MN, B_MN, EPI_F32>(ta, tb, ta2, p, split_k, batch, st);
}

int g_force_bn = 0;

}  // namespace

// shared with gemm_grouped.cu (hidden visibility: internal to the library)
int pcm_get_tensor_map_2d(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                          CUtensorMap* out) {
    return get_tensor_map(ptr, inner, outer, ld, box_inner, box_outer, out);
}

// Debug aid for tile-shape sweeps (tools/bench_gemm.py): force the N extent of the output tile of
// subsequent GEMM launches (64 / 128 / 256; 0 = heuristic).
PCM_API int pcm_gemm_debug_force_bn(int bn) {
    g_force_bn = bn;
    return PCM_OK;
}

// Extended entry point: batched, scaled, with head-split / head-merge output addressing (used by
// the attention GEMMs).  Operand tensor maps span `batch` stacked problems: a_rows_total /
// b_rows_total are the row counts of the full 2-D operand tensors.
PCM_API int pcm_gemm_bf16_ex(int M, int N, int K, int batch, const void* A, int lda, int a_mn, long long a_rows_total,
                             long long a_batch_rows, const void* B, int ldb, int b_mn, long long b_rows_total,
                             long long b_batch_rows, void* C, int ldc, int c_bf16, int c_mode, long long c_batch_rows,
                             int hs_B, int hs_nh, int hs_L, float alpha, const float* bias, int relu, int accumulate,
                             int split_k, pcm_stream_t stream) {
    return pcm_gemm_bf16_ex2(M, N, K, batch, A, lda, a_mn, a_rows_total, a_batch_rows, B, ldb, b_mn, b_rows_total, b_batch_rows, C,
                             ldc, c_bf16, c_mode, c_batch_rows, hs_B, hs_nh, hs_L, alpha, bias, relu, accumulate, split_k,
                             nullptr, 0, 0, 0, 0, stream);
}

// _ex2: + second A operand for output columns >= a2_from_col (same shape / majorness / batching as A, row pitch lda2)
// and head-split output in parts of hs_part_cols columns, hs_part_stride elements apart (see GemmParams).
PCM_API int pcm_gemm_bf16_ex2(int M, int N, int K, int batch, const void* A, int lda, int a_mn, long long a_rows_total,
                              long long a_batch_rows, const void* B, int ldb, int b_mn, long long b_rows_total,
                              long long b_batch_rows, void* C, int ldc, int c_bf16, int c_mode, long long c_batch_rows,
                              int hs_B, int hs_nh, int hs_L, float alpha, const float* bias, int relu, int accumulate,
                              int split_k, const void* A2, int lda2, int a2_from_col, int hs_part_cols,
                              long long hs_part_stride, pcm_stream_t stream) {
    if (M <= 0 || N <= 0 || batch <= 0) return PCM_OK;
    if (!A || !B || !C || K <= 0) return PCM_EINVAL;
    if (A2 && ((lda2 % 8) || (reinterpret_cast<uintptr_t>(A2) & 15) || a2_from_col <= 0)) return PCM_EUNSUPPORTED;
    if (hs_part_cols < 0 || (hs_part_cols > 0 && (c_mode != 1 || (hs_part_cols % 64)))) return PCM_EINVAL;
    if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return PCM_EUNSUPPORTED;  // TMA: 16-byte aligned base and row pitch
    if ((split_k > 1 || accumulate) && c_bf16) return PCM_EINVAL;
    if (split_k > 1 && (bias || relu)) return PCM_EINVAL;
    if (c_mode < 0 || c_mode > 2 || (c_mode != 0 && (hs_B <= 0 || hs_nh <= 0))) return PCM_EINVAL;
    if (c_mode == 1 && (N % 64)) return PCM_EUNSUPPORTED;
    const int kblocks = (K + BLOCK_K - 1) / BLOCK_K;
    // wide tiles cut L2->smem operand traffic per FLOP (ncu: 128x128 SS tiles are smem/L2 bound)
    int BN = (N <= 64) ? 64 : ((N >= 256 && ((N + 255) / 256 * 256 - N) < 128) ? 256 : 128);
    if (split_k <= 0) {
        // auto (weight-gradient GEMMs: small output, long K): tile width and K split chosen together
        // from the measured sweep (tools/bench_gemm_sweep.py): short K wants 64-wide tiles and few
        // slices (every slice costs a tile of fp32 atomics), long K with one row of tiles wants
        // 128-wide tiles, otherwise wide tiles; slices fill one wave of the SMs, >= 8 k-blocks each
        if (!accumulate) return PCM_EINVAL;
        const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
        if (kblocks <= 128) BN = 64;
        else if (m_tiles == 1 && N >= 128) BN = 128;
        const int tiles = m_tiles * ((N + BN - 1) / BN);
        int want = 148 / (tiles > 0 ? tiles : 1);
        if (want < 1) want = 1;
        int cap = kblocks / 8;
        if (cap < 1) cap = 1;
        split_k = want < cap ? want : cap;
    }
    if (g_force_bn == 64 || g_force_bn == 128 || g_force_bn == 256) BN = g_force_bn;  // tools/bench_gemm_sweep.py
    while (A2 && BN > 64 && (a2_from_col % BN)) BN >>= 1;  // the operand switch happens on output-tile boundaries
    if (split_k > kblocks) split_k = kblocks;
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.k_blocks_per_split = (kblocks + split_k - 1) / split_k;
    split_k = (kblocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
    p.split_k = split_k;
    p.batch = batch;
    p.a_batch_rows = a_batch_rows; p.b_batch_rows = b_batch_rows;
    p.C = C; p.ldc = ldc; p.c_mode = c_mode; p.c_batch_rows = c_batch_rows;
    p.hs_B = hs_B; p.hs_nh = hs_nh; p.hs_L = hs_L;
    p.bias = bias; p.relu = relu; p.c_bf16 = c_bf16; p.alpha = alpha;
    p.atomic = (accumulate || split_k > 1) ? 1 : 0;
    p.a2_from_col = 0x7fffffff;
    p.hs_part_cols = hs_part_cols;
    p.hs_part_stride = hs_part_stride;
    CUtensorMap ta, tb, ta2;
    int r;
    if (!a_mn) r = get_tensor_map(A, (uint64_t)K, (uint64_t)a_rows_total, (uint64_t)lda, 64, BLOCK_M, &ta);
    else r = get_tensor_map(A, (uint64_t)M, (uint64_t)a_rows_total, (uint64_t)lda, 64, BLOCK_K, &ta);
    if (r) return r;
    ta2 = ta;
    if (A2) {
        if (a2_from_col % BN) return PCM_EUNSUPPORTED;  // the switch happens on output-tile boundaries
        if (!a_mn) r = get_tensor_map(A2, (uint64_t)K, (uint64_t)a_rows_total, (uint64_t)lda2, 64, BLOCK_M, &ta2);
        else r = get_tensor_map(A2, (uint64_t)M, (uint64_t)a_rows_total, (uint64_t)lda2, 64, BLOCK_K, &ta2);
        if (r) return r;
        p.a2_from_col = a2_from_col;
    }
    if (!b_mn) r = get_tensor_map(B, (uint64_t)K, (uint64_t)b_rows_total, (uint64_t)ldb, 64, BN, &tb);
    else r = get_tensor_map(B, (uint64_t)N, (uint64_t)b_rows_total, (uint64_t)ldb, 64, BLOCK_K, &tb);
    if (r) return r;
    cudaStream_t st = pcm_cu_stream(stream);
    if (BN == 64) {
        if (!a_mn && !b_mn) return launch<64, false, false>(ta, tb, ta2, p, split_k, batch, st);
        if (!a_mn && b_mn) return launch<64, false, true>(ta, tb, ta2, p, split_k, batch, st);
        if (a_mn && !b_mn) return launch<64, true, false>(ta, tb, ta2, p, split_k, batch, st);
        return launch<64, true, true>(ta, tb, ta2, p, split_k, batch, st);
    }
    if (BN == 256) {
        if (!a_mn && !b_mn) return launch<256, false, false>(ta, tb, ta2, p, split_k, batch, st);
        if (!a_mn && b_mn) return launch<256, false, true>(ta, tb, ta2, p, split_k, batch, st);
        if (a_mn && !b_mn) return launch<256, true, false>(ta, tb, ta2, p, split_k, batch, st);
        return launch<256, true, true>(ta, tb, ta2, p, split_k, batch, st);
    }
    if (!a_mn && !b_mn) return launch<128, false, false>(ta, tb, ta2, p, split_k, batch, st);
    if (!a_mn && b_mn) return launch<128, false, true>(ta, tb, ta2, p, split_k, batch, st);
    if (a_mn && !b_mn) return launch<128, true, false>(ta, tb, ta2, p, split_k, batch, st);
    return launch<128, true, true>(ta, tb, ta2, p, split_k, batch, st);
}

// C[m, n] (+)= sum_k A(m, k) B(n, k) (+ bias[n]) (ReLU).  a_mn / b_mn = 0: operand stored row-major
// [rows, K] with pitch ld (K contiguous); = 1: stored row-major [K, rows] with pitch ld (rows
// contiguous).  c_bf16: output dtype; accumulate: atomically add into fp32 C (required if split_k > 1).
PCM_API int pcm_gemm_bf16(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn,
                          void* C, int ldc, int c_bf16, const float* bias, int relu, int accumulate, int split_k,
                          pcm_stream_t stream) {
    return pcm_gemm_bf16_ex(M, N, K, 1, A, lda, a_mn, a_mn ? K : M, 0, B, ldb, b_mn, b_mn ? K : N, 0, C, ldc, c_bf16, 0, 0,
                            0, 0, 0, 1.0f, bias, relu, accumulate, split_k, stream);
}
