// gemm_grouped.cu -- grouped weight-gradient GEMM on tcgen05: many independent problems
//
//     C_p[m, n] += sum_r dY_p[r, m] * X_p[r, n]          (dW = dY^T X, fp32 accumulation into the flat gradient)
//
// in ONE persistent launch.  The training step forms ~110 weight gradients; two thirds of them belong to the decoder /
// CVAE-encoder layers whose 6 400-row problems (3.4 GFLOP each) are launch- and tail-latency-bound when launched one by
// one (12 us per launch at 280 TFLOP/s, 80 launches per step).  They do not depend on each other and nothing in the
// backward chain waits for them, so the operator layer queues them (functional._dw) and this kernel runs the queue as one
// flattened list of (problem, k-slice, m-block, n-block) work items over all SMs.
//
// Same machine mapping as gemm_tcgen05.cu (warp 0 TMA producer, warp 1 MMA issuer, 8 epilogue warps, 2 TMEM accumulators,
// SWIZZLE_128B ring), specialised to MN-major A and B (both operands are read in the layout the activations already
// have) and the red.global.add.f32 epilogue.  Per-problem tensor maps and shapes travel in the kernel parameter block
// (__grid_constant__, ~17 KB for 56 problems): nothing is staged through device memory, so a CUDA-graph capture of the
// launch is self-contained.
#include "common.cuh"
#include "tcgen05_ptx.cuh"

#include <cuda.h>

int pcm_get_tensor_map_2d(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                          CUtensorMap* out);  // gemm_tcgen05.cu (cached encoder)

namespace {

using namespace pcm_tc;

constexpr int GP_MAX = 56;  // problems per launch (parameter block: 56 x (2 x 128 B maps + 40 B) = 16.6 KB)
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;
constexpr int NUM_EPI_WARPS = 8;
template <int BLOCK_N> struct StagesFor { static constexpr int value = BLOCK_N == 256 ? 4 : 6; };

struct GroupedProblem {
    float* C;
    int ldc, M, N, kblocks;
    int kbps;      // k-blocks per slice
    int tiles_n;   // output tiles along N
    int tiles_mn;  // output tiles per k-slice
};
struct GroupedParams {
    int n;
    int item_prefix[GP_MAX + 1];  // work items before problem p
    GroupedProblem prob[GP_MAX];
};
struct GroupedMaps {
    CUtensorMap ta[GP_MAX], tb[GP_MAX];
};

struct Item { int p, m_blk, n_blk, kb0, kb1; };

__device__ __forceinline__ Item decode(const GroupedParams& g, int t) {
    int lo = 0, hi = g.n - 1;
    while (lo < hi) {  // last problem whose prefix <= t
        const int mid = (lo + hi + 1) >> 1;
        if (g.item_prefix[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const GroupedProblem& q = g.prob[lo];
    const int local = t - g.item_prefix[lo];
    const int ks = local / q.tiles_mn, r = local - ks * q.tiles_mn;
    Item it;
    it.p = lo;
    it.m_blk = r / q.tiles_n;
    it.n_blk = r - it.m_blk * q.tiles_n;
    it.kb0 = ks * q.kbps;
    it.kb1 = min(q.kblocks, it.kb0 + q.kbps);
    return it;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_dw_grouped_kernel(const __grid_constant__ GroupedMaps maps,
                                                                         const __grid_constant__ GroupedParams g) {
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int STAGES = StagesFor<BLOCK_N>::value;
    constexpr uint32_t TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    // D = F32, A = B = BF16, both operands MN-major
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint8_t* epi_stage = tiles + STAGES * STAGE_BYTES + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_items = g.item_prefix[g.n];

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();

    if (warp == 0) {
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
            const Item w = decode(g, t);
            const CUtensorMap* ta = &maps.ta[w.p];
            const CUtensorMap* tb = &maps.tb[w.p];
            for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (elect_one_sync()) {
                    uint8_t* sa = tiles + s * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    const int k0 = kb * BLOCK_K;
#pragma unroll
                    for (int c = 0; c < BLOCK_M / 64; ++c)
                        tma_load_2d(sa + c * (BLOCK_K * 128), ta, &full_bar[s], w.m_blk * BLOCK_M + c * 64, k0);
#pragma unroll
                    for (int c = 0; c < BLOCK_N / 64; ++c)
                        tma_load_2d(sb + c * (BLOCK_K * 128), tb, &full_bar[s], w.n_blk * BLOCK_N + c * 64, k0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        uint32_t it = 0;
        int i = 0;
        for (int t = blockIdx.x; t < total_items; t += gridDim.x, ++i) {
            const Item w = decode(g, t);
            const int nkb = w.kb1 - w.kb0;
            const int acc = i & 1;
            mbar_wait(&tmem_empty_bar[acc], (((uint32_t)i >> 1) & 1) ^ 1);
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
                        const uint64_t da = make_smem_desc(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
                        const uint64_t db = make_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
                        umma_f16(tmem_d, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int chalf = (warp - 2) >> 2;
        uint8_t* stage = epi_stage + (warp - 2) * 4096;
        constexpr int NLD = BLOCK_N / 64 > 0 ? BLOCK_N / 64 : 1;
        int i = 0;
        for (int t = blockIdx.x; t < total_items; t += gridDim.x, ++i) {
            const Item w = decode(g, t);
            const GroupedProblem& q = g.prob[w.p];
            const int acc = i & 1;
            const int row_base = w.m_blk * BLOCK_M + quad * 32;
            const int cbase = chalf * (BLOCK_N / 2);
            const int col_lim = min(q.N, w.n_blk * BLOCK_N + cbase + BLOCK_N / 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + cbase);
            const size_t row_dst = (size_t)(row_base + lane) * q.ldc;
            mbar_wait(&tmem_full_bar[acc], ((uint32_t)i >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
#pragma unroll 1
            for (int ld = 0; ld < NLD; ++ld) {
                const int col0 = w.n_blk * BLOCK_N + cbase + ld * 32;
                tmem_ld_32x32b_x32(taddr + (uint32_t)(ld * 32), v);
                tmem_ld_wait(v);
                if (col0 >= q.N) continue;  // warp-uniform
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                        make_float4(__uint_as_float(v[c * 4 + 0]), __uint_as_float(v[c * 4 + 1]), __uint_as_float(v[c * 4 + 2]),
                                    __uint_as_float(v[c * 4 + 3]));
                __syncwarp();
#pragma unroll 4
                for (int rl = 0; rl < 32; ++rl) {
                    const size_t rd = __shfl_sync(PCM_FULL_MASK, row_dst, rl);
                    const float x = *reinterpret_cast<const float*>(stage + rl * 128 + (((lane >> 2) ^ (rl & 7)) << 4) + ((lane & 3) << 2));
                    if (row_base + rl < q.M && col0 + lane < col_lim) atomicAdd(q.C + rd + col0 + lane, x);
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BLOCK_N>
int launch_group(const GroupedMaps& maps, const GroupedParams& g, cudaStream_t st) {
    constexpr size_t SMEM = StagesFor<BLOCK_N>::value * (BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2) + 1024 + 256 + NUM_EPI_WARPS * 4096;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_dw_grouped_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
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
    const int items = g.item_prefix[g.n];
    const int grid = items < num_sms ? items : num_sms;
    cudaError_t le = pcm_launch(gemm_dw_grouped_kernel<BLOCK_N>, dim3(grid), dim3(NUM_THREADS), SMEM, st, maps, g);
    if (le != cudaSuccess) return (int)le;
    return pcm_launch_status();
}

}  // namespace

// n problems  C_p (M_p x N_p fp32, row pitch ldc_p) += A_p^T B_p  with A_p = (K_p x M_p) bf16 (row pitch lda_p) and
// B_p = (K_p x N_p) bf16 (row pitch ldb_p): the weight-gradient form dW = dY^T X with both operands read in place.
// All arrays are HOST arrays of length n.  Problems are bucketed by tile width (N <= 64: 128 x 64 tiles, else 128 x 256)
// and launched in chunks of at most 56; k is sliced so that no work item exceeds 128 k-blocks and one launch offers at
// least two waves of work items.
PCM_API int pcm_gemm_dw_grouped(int n, const void* const* A, const int* lda, const void* const* B, const int* ldb, float* const* C,
                                const int* ldc, const int* M, const int* N, const int* K, pcm_stream_t stream) {
    if (n <= 0) return PCM_OK;
    if (!A || !lda || !B || !ldb || !C || !ldc || !M || !N || !K) return PCM_EINVAL;
    cudaStream_t st = pcm_cu_stream(stream);
    for (int cls = 0; cls < 2; ++cls) {
        const int BN = cls == 0 ? 64 : 256;
        int idx[GP_MAX];
        int cnt = 0;
        auto flush = [&]() -> int {
            if (cnt == 0) return PCM_OK;
            GroupedMaps maps;
            GroupedParams g;
            g.n = cnt;
            long base_items = 0;
            for (int j = 0; j < cnt; ++j) {
                const int p = idx[j];
                GroupedProblem& q = g.prob[j];
                q.C = C[p]; q.ldc = ldc[p]; q.M = M[p]; q.N = N[p];
                q.kblocks = (K[p] + BLOCK_K - 1) / BLOCK_K;
                q.tiles_n = (N[p] + BN - 1) / BN;
                q.tiles_mn = ((M[p] + BLOCK_M - 1) / BLOCK_M) * q.tiles_n;
                int split = (q.kblocks + 127) / 128;
                q.kbps = (q.kblocks + split - 1) / split;
                base_items += (long)q.tiles_mn * ((q.kblocks + q.kbps - 1) / q.kbps);
                int r = pcm_get_tensor_map_2d(A[p], (uint64_t)M[p], (uint64_t)K[p], (uint64_t)lda[p], 64, BLOCK_K, &maps.ta[j]);
                if (r) return r;
                r = pcm_get_tensor_map_2d(B[p], (uint64_t)N[p], (uint64_t)K[p], (uint64_t)ldb[p], 64, BLOCK_K, &maps.tb[j]);
                if (r) return r;
            }
            if (base_items < 296) {  // too few items for two waves: slice k finer (>= 8 k-blocks per item)
                const int mult = (int)((296 + base_items - 1) / base_items);
                for (int j = 0; j < cnt; ++j) {
                    GroupedProblem& q = g.prob[j];
                    int kbps = q.kbps / mult;
                    if (kbps < 8) kbps = q.kbps < 8 ? q.kbps : 8;
                    q.kbps = kbps;
                }
            }
            int items = 0;
            for (int j = 0; j < cnt; ++j) {
                g.item_prefix[j] = items;
                items += g.prob[j].tiles_mn * ((g.prob[j].kblocks + g.prob[j].kbps - 1) / g.prob[j].kbps);
            }
            g.item_prefix[cnt] = items;
            for (int j = cnt + 1; j <= GP_MAX; ++j) g.item_prefix[j] = items;
            cnt = 0;
            return BN == 64 ? launch_group<64>(maps, g, st) : launch_group<256>(maps, g, st);
        };
        for (int p = 0; p < n; ++p) {
            if (M[p] <= 0 || N[p] <= 0 || K[p] <= 0) continue;
            if (!A[p] || !B[p] || !C[p]) return PCM_EINVAL;
            if ((lda[p] % 8) || (ldb[p] % 8) || (reinterpret_cast<uintptr_t>(A[p]) & 15) || (reinterpret_cast<uintptr_t>(B[p]) & 15))
                return PCM_EUNSUPPORTED;
            const bool narrow = N[p] <= 64;
            if (narrow != (cls == 0)) continue;
            idx[cnt++] = p;
            if (cnt == GP_MAX) {
                const int r = flush();
                if (r) return r;
            }
        }
        const int r = flush();
        if (r) return r;
    }
    return PCM_OK;
}
