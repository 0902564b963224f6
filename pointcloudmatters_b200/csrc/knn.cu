// knn.cu -- kNN / ball / random-ball query for sm_100a.
//
// kNN replaces knn_query_cuda_kernel (reference libs/pointops/src/knn_query/
// knn_query_cuda_kernel.cu:60-104).  The reference's output ORDER under equal distances is an
// artefact of its sequential binary max-heap (strict `d2 < root` replacement, `reheap`, then
// `heap_sort`), so the only way to be bit-exact on every input is to replay that heap -- for the
// queries that HAVE ties.  Two kernels: knn_warp_kernel (nsample <= 31, the default: one warp per
// query, sorted per-lane list, heap replay only for queries with a tie among their k + 1 smallest;
// see its comment below) and knn_kernel (any nsample <= 128: always replays the heap).  The latter
// keeps one query per thread like the reference, but
//   * the heap lives in SHARED memory laid out [slot][thread] (bank = thread id: conflict-free
//     for any per-thread slot), not in a 1 KB/thread local-memory stack frame;
//   * the cloud is staged through shared memory in float4 tiles with coalesced loads, so the
//     inner loop is one broadcast LDS.128 + 6 FP ops per candidate instead of three global loads;
//   * the heap root is cached in a register, so the common reject path touches no memory;
//   * results leave through a block-wide transposed, fully coalesced store.
// Algorithmic traffic: 12N + 12M + 8*M*k bytes per cloud.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int KNN_T = 64;       // queries per CTA (small CTAs: many co-resident, >= 16 warps/SM)
constexpr int KNN_TILE = 1024;  // source points staged per tile (float4 -> 16 KB)

// Sift (nd, ni) down from the root of a max-heap of `size` slots.  "Hole" formulation of the
// reference's swap loop (knn_query_cuda_kernel.cu:15-30): identical comparisons, identical result.
__device__ __forceinline__ void heap_sift(float* hd, int* hi, int stride, int size, float nd, int ni) {
    int pos = 0;
    int child = 1;
    while (child < size) {
        float cd = hd[child * stride];
        if (child + 1 < size) {
            const float rd = hd[(child + 1) * stride];
            if (rd > cd) { cd = rd; child++; }
        }
        if (nd > cd) break;
        hd[pos * stride] = cd;
        hi[pos * stride] = hi[child * stride];
        pos = child;
        child = pos * 2 + 1;
    }
    hd[pos * stride] = nd;
    hi[pos * stride] = ni;
}

// heap_sort (knn_query_cuda_kernel.cu:33-42): swap(0, i) then reheap over the first i slots.
__device__ __forceinline__ void heap_sort(float* hd, int* hi, int stride, int k) {
    for (int i = k - 1; i > 0; i--) {
        const float nd = hd[i * stride];
        const int ni = hi[i * stride];
        hd[i * stride] = hd[0];
        hi[i * stride] = hi[0];
        heap_sift(hd, hi, stride, i, nd, ni);
    }
}

__global__ void __launch_bounds__(KNN_T) knn_kernel(int b, int m, int k,
                                                    const float* __restrict__ xyz,
                                                    const float* __restrict__ new_xyz,
                                                    const int* __restrict__ offset,
                                                    const int* __restrict__ new_offset,
                                                    int* __restrict__ idx, float* __restrict__ dist2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* hd_all = reinterpret_cast<float*>(smem_raw);
    int* hi_all = reinterpret_cast<int*>(hd_all + (size_t)k * KNN_T);
    float4* tile = reinterpret_cast<float4*>(hi_all + (size_t)k * KNN_T);
    float* tile_f = reinterpret_cast<float*>(tile);

    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * KNN_T;
    const int q = q0 + tid;
    const bool active = q < m;
    const int q_last = min(m, q0 + KNN_T) - 1;
    const int c_first = pcm_cloud_of(q0, new_offset, b);
    const int c_last = pcm_cloud_of(q_last, new_offset, b);
    const int my_cloud = active ? pcm_cloud_of(q, new_offset, b) : -1;

    float* hd = hd_all + tid;
    int* hi = hi_all + tid;
    for (int i = 0; i < k; ++i) { hd[i * KNN_T] = 1e10f; hi[i * KNN_T] = -1; }
    float root = 1e10f;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = new_xyz[(size_t)q * 3 + 0]; qy = new_xyz[(size_t)q * 3 + 1]; qz = new_xyz[(size_t)q * 3 + 2]; }

    for (int c = c_first; c <= c_last; ++c) {
        const int s = c ? __ldg(offset + c - 1) : 0;
        const int e = __ldg(offset + c);
        for (int t0 = s; t0 < e; t0 += KNN_TILE) {
            const int cnt = min(KNN_TILE, e - t0);
            __syncthreads();
            const float* src = xyz + (size_t)t0 * 3;
            for (int f = tid; f < cnt * 3; f += KNN_T) {
                const int p = f / 3;
                tile_f[p * 4 + (f - p * 3)] = __ldg(src + f);
            }
            __syncthreads();
            if (my_cloud == c) {
                // 4 candidates per trip: the distances are independent (ILP), and the heap is
                // only entered -- strictly in scan order, so the replay stays exact -- when one
                // of them beats the cached root.
                int i = 0;
                for (; i + 4 <= cnt; i += 4) {
                    float d[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 P = tile[i + u];
                        d[u] = pcm_dist2(qx - P.x, qy - P.y, qz - P.z);
                    }
                    if (fminf(fminf(d[0], d[1]), fminf(d[2], d[3])) < root) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (d[u] < root) {
                                heap_sift(hd, hi, KNN_T, k, d[u], t0 + i + u);
                                root = hd[0];
                            }
                        }
                    }
                }
                for (; i < cnt; ++i) {
                    const float4 P = tile[i];
                    const float d2 = pcm_dist2(qx - P.x, qy - P.y, qz - P.z);
                    if (d2 < root) {
                        heap_sift(hd, hi, KNN_T, k, d2, t0 + i);
                        root = hd[0];
                    }
                }
            }
        }
    }
    heap_sort(hd, hi, KNN_T, k);
    __syncthreads();
    const int nq = q_last - q0 + 1;
    const int total = nq * k;
    int* o_idx = idx + (size_t)q0 * k;
    float* o_d = dist2 ? dist2 + (size_t)q0 * k : nullptr;
    for (int j = tid; j < total; j += KNN_T) {
        const int ql = j / k;
        const int i = j - ql * k;
        o_idx[j] = hi_all[i * KNN_T + ql];
        if (o_d) o_d[j] = hd_all[i * KNN_T + ql];
    }
}

// ------------------------------------------------------------------------------------------
// kNN, one WARP per query (nsample <= 31).
//
// The thread-per-query kernel above spends most of its issue slots in divergent heap sifts (one lane at a time, ~40
// instructions each, ~k (1 + ln(N / k)) of them per query) and keeps 10 % of the SM's warp slots busy.  Here the 32 lanes
// of a warp scan 32 candidates per step for one query, and the running result is a SORTED list with one entry per lane
// (lane i = (i+1)-th smallest so far): a candidate that beats the list's threshold is inserted with two shuffles, a ballot
// and a select -- ~8 warp instructions.
//
// Exactness.  The reference's output is its max-heap after heap_sort.  If the k + 1 smallest distances of a query are
// pairwise different, the heap's final CONTENT is the unique set of the k smallest and heap_sort's output is that set in
// ascending order, whatever the heap's internal arrangement was: the sorted list IS the reference result.  The list keeps
// k + 1 entries for exactly this check (candidates equal to its threshold cannot change ranks 1..k and are skipped).  A
// query that shows a tie among its k + 1 smallest is REPLAYED through the reference's heap (phase 2): an element enters
// the reference heap iff its distance is below the k-th smallest of the prefix before it (the heap root), a condition on
// VALUES only, so the warp finds the entering elements with the same 32-wide scan (ballot against the root), feeds them
// to heap_sift in scan order (lane 0, shared-memory heap), and finishes with the reference's heap_sort.
// ------------------------------------------------------------------------------------------
constexpr int KW_WARPS = 8;
constexpr int KW_TILE = 1024;  // source points staged per tile (float4 -> 16 KB)

// insert (cd, ci) into the ascending per-lane list (bd, bi): every lane decides on its own from its entry and its left
// neighbour's (ud, ui) -- entries <= cd stay, the first greater one takes the candidate, the rest move up by one lane
__device__ __forceinline__ void list_insert(float& bd, int& bi, float cd, int ci, int lane) {
    float ud = __shfl_up_sync(PCM_FULL_MASK, bd, 1);
    const int ui = __shfl_up_sync(PCM_FULL_MASK, bi, 1);
    if (lane == 0) ud = -INFINITY;
    const bool stay = bd <= cd, first = ud <= cd;
    bi = stay ? bi : (first ? ci : ui);
    bd = stay ? bd : fmaxf(cd, ud);
}

template <int QPW>
__global__ void __launch_bounds__(KW_WARPS * 32) knn_warp_kernel(int b, int m, int k, const float* __restrict__ xyz,
                                                                  const float* __restrict__ new_xyz,
                                                                  const int* __restrict__ offset,
                                                                  const int* __restrict__ new_offset, int* __restrict__ idx,
                                                                  float* __restrict__ dist2) {
    __shared__ float4 tile[KW_TILE];
    __shared__ float s_hd[KW_WARPS][QPW][32];
    __shared__ int s_hi[KW_WARPS][QPW][32];
    float* tile_f = reinterpret_cast<float*>(tile);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int QPC = KW_WARPS * QPW;
    const int q0 = blockIdx.x * QPC;
    const int q_last = min(m, q0 + QPC) - 1;
    const int c_first = pcm_cloud_of(q0, new_offset, b);
    const int c_last = pcm_cloud_of(q_last, new_offset, b);

    float qx[QPW], qy[QPW], qz[QPW], bd[QPW], thr[QPW], thr1[QPW];  // thr = entry k (the k+1-th smallest), thr1 = entry k-1
    int bi[QPW], qc[QPW];
#pragma unroll
    for (int j = 0; j < QPW; ++j) {
        const int q = q0 + warp * QPW + j;
        const bool act = q < m;
        qc[j] = act ? pcm_cloud_of(q, new_offset, b) : -1;
        qx[j] = act ? new_xyz[(size_t)q * 3 + 0] : 0.f;
        qy[j] = act ? new_xyz[(size_t)q * 3 + 1] : 0.f;
        qz[j] = act ? new_xyz[(size_t)q * 3 + 2] : 0.f;
        bd[j] = 1e10f; bi[j] = -1; thr[j] = 1e10f; thr1[j] = 1e10f;
    }
    unsigned replay = 0;  // bit j: query j of this warp goes through the exact heap replay

    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1) {
#pragma unroll
            for (int j = 0; j < QPW; ++j) {
                s_hd[warp][j][lane] = 1e10f; s_hi[warp][j][lane] = -1;
                thr[j] = 1e10f;  // phase 2: the heap root
            }
            __syncwarp();
        }
        for (int c = c_first; c <= c_last; ++c) {
            const int s = c ? __ldg(offset + c - 1) : 0;
            const int e = __ldg(offset + c);
            unsigned mine = 0;  // queries of this warp that scan cloud c in this phase (warp-uniform)
#pragma unroll
            for (int j = 0; j < QPW; ++j)
                if (qc[j] == c && (phase == 0 || ((replay >> j) & 1u))) mine |= 1u << j;
            for (int t0 = s; t0 < e; t0 += KW_TILE) {
                const int cnt = min(KW_TILE, e - t0);
                __syncthreads();
                const float* src = xyz + (size_t)t0 * 3;
                for (int f = tid; f < cnt * 3; f += KW_WARPS * 32) {
                    const int pp = f / 3;
                    tile_f[pp * 4 + (f - pp * 3)] = __ldg(src + f);
                }
                __syncthreads();
                if (!mine) continue;
                for (int i0 = 0; i0 < cnt; i0 += 32) {
                    const bool valid = i0 + lane < cnt;
                    const float4 P = tile[valid ? i0 + lane : i0];
#pragma unroll
                    for (int j = 0; j < QPW; ++j) {
                        if (!((mine >> j) & 1u)) continue;
                        const float d = pcm_dist2(qx[j] - P.x, qy[j] - P.y, qz[j] - P.z);
                        unsigned mk = __ballot_sync(PCM_FULL_MASK, valid && d < thr[j]);
                        while (mk) {  // candidates in scan order; the threshold only falls, so each is re-tested
                            const int sl = __ffs(mk) - 1;
                            mk &= mk - 1;
                            const float cd = __shfl_sync(PCM_FULL_MASK, d, sl);
                            if (cd < thr[j]) {
                                if (phase == 0) {
                                    list_insert(bd[j], bi[j], cd, t0 + i0 + sl, lane);
                                    thr[j] = fmaxf(cd, thr1[j]);  // cd < old entry k: the new entry k is max(cd, old entry k-1)
                                    thr1[j] = __shfl_sync(PCM_FULL_MASK, bd[j], k - 1);
                                } else {
                                    // only lane 0 touches the heap; the new root travels by shuffle (which also reconverges
                                    // the warp), so no lane reads shared memory that lane 0 is about to rewrite
                                    float root = 0.f;
                                    if (lane == 0) {
                                        heap_sift(s_hd[warp][j], s_hi[warp][j], 1, k, cd, t0 + i0 + sl);
                                        root = s_hd[warp][j][0];
                                    }
                                    thr[j] = __shfl_sync(PCM_FULL_MASK, root, 0);
                                }
                            }
                        }
                    }
                }
            }
        }
        if (phase == 0) {
            // ties among the k + 1 smallest (entries of untouched slots, idx -1, are identical and harmless)
#pragma unroll
            for (int j = 0; j < QPW; ++j) {
                const float nd = __shfl_down_sync(PCM_FULL_MASK, bd[j], 1);
                const int ni = __shfl_down_sync(PCM_FULL_MASK, bi[j], 1);
                const bool tie = lane < k && bd[j] == nd && !(bi[j] == -1 && ni == -1);
                if (__ballot_sync(PCM_FULL_MASK, tie) != 0u && qc[j] >= 0) replay |= 1u << j;
            }
            if (!__syncthreads_or(replay != 0u)) break;  // CTA-uniform: the tile staging of phase 2 needs every warp
        } else {
#pragma unroll
            for (int j = 0; j < QPW; ++j) {
                if (!((replay >> j) & 1u)) continue;
                if (lane == 0) heap_sort(s_hd[warp][j], s_hi[warp][j], 1, k);
                __syncwarp();
                bd[j] = s_hd[warp][j][lane];
                bi[j] = s_hi[warp][j][lane];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < QPW; ++j) {
        const int q = q0 + warp * QPW + j;
        if (q < m && lane < k) {
            idx[(size_t)q * k + lane] = bi[j];
            if (dist2) dist2[(size_t)q * k + lane] = bd[j];
        }
    }
}

template <int QPW>
static int launch_knn_warp(int b, int m, int k, const float* xyz, const float* new_xyz, const int* offset,
                           const int* new_offset, int* idx, float* dist2, cudaStream_t st) {
    knn_warp_kernel<QPW><<<pcm_divup(m, KW_WARPS * QPW), KW_WARPS * 32, 0, st>>>(b, m, k, xyz, new_xyz, offset, new_offset, idx,
                                                                              dist2);
    return pcm_launch_status();
}

// ------------------------------------------------------------------------------------------
// Ball query: replaces ball_query_cuda_kernel (ball_query_cuda_kernel.cu:58-123).
// One WARP per query: lanes scan the cloud 32 points at a time and compact the hits IN SCAN
// ORDER (ballot + popc prefix) into a per-warp shared-memory list, replacing the reference's
// 16 KB-per-thread local arrays.  Lane 0 then replays the reference's heap_sort on that list
// (applied, as in the reference, to an array that was never heapified) and the warp writes the
// result.  Collection stops at 2048 candidates (the reference overflows its arrays there).
// ------------------------------------------------------------------------------------------
constexpr int BALL_MAX = 2048;
constexpr int BALL_WARPS = 4;

__device__ __forceinline__ bool ball_hit(float d2, float min_r2, float max_r2) {
    return ((double)d2 <= 1e-5) || (d2 >= min_r2 && d2 < max_r2);
}

__global__ void __launch_bounds__(BALL_WARPS * 32) ball_query_kernel(
    int b, int m, int nsample, float min_radius, float max_radius, const float* __restrict__ xyz,
    const float* __restrict__ new_xyz, const int* __restrict__ offset,
    const int* __restrict__ new_offset, int* __restrict__ idx, float* __restrict__ dist2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* cd = reinterpret_cast<float*>(smem_raw) + (size_t)warp * BALL_MAX;
    int* ci = reinterpret_cast<int*>(smem_raw + (size_t)BALL_WARPS * BALL_MAX * sizeof(float)) + (size_t)warp * BALL_MAX;

    const int pt = blockIdx.x * BALL_WARPS + warp;
    if (pt >= m) return;
    const int bt = pcm_cloud_of(pt, new_offset, b);
    const int start = bt ? __ldg(offset + bt - 1) : 0;
    const int end = __ldg(offset + bt);
    const float max_r2 = __fmul_rn(max_radius, max_radius);
    const float min_r2 = __fmul_rn(min_radius, min_radius);
    const float qx = new_xyz[(size_t)pt * 3 + 0], qy = new_xyz[(size_t)pt * 3 + 1], qz = new_xyz[(size_t)pt * 3 + 2];

    int num = 0;
    for (int base = start; base < end && num < BALL_MAX; base += 32) {
        const int i = base + lane;
        bool hit = false;
        float d2 = 0.f;
        if (i < end) {
            const float* p = xyz + (size_t)i * 3;
            d2 = pcm_dist2(qx - p[0], qy - p[1], qz - p[2]);
            hit = ball_hit(d2, min_r2, max_r2);
        }
        const unsigned mask = __ballot_sync(PCM_FULL_MASK, hit);
        const int pos = num + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < BALL_MAX) { cd[pos] = d2; ci[pos] = i; }
        num += __popc(mask);
    }
    if (num > BALL_MAX) num = BALL_MAX;
    __syncwarp();
    if (lane == 0) heap_sort(cd, ci, 1, num);
    __syncwarp();
    int* o_idx = idx + (size_t)pt * nsample;
    float* o_d = dist2 + (size_t)pt * nsample;
    if (num <= nsample) {
        for (int i = lane; i < nsample; i += 32) {
            o_idx[i] = i < num ? ci[i] : -1;
            o_d[i] = i < num ? cd[i] : 1e10f;
        }
    } else {
        const float sep = __fdiv_rn((float)num, (float)nsample);
        for (int i = lane; i < nsample; i += 32) {
            const int index = (int)__fmul_rn(sep, (float)i);
            o_idx[i] = ci[index];
            o_d[i] = (float)ci[index];  // reference ball_query_cuda_kernel.cu:120 (sic)
        }
    }
}

// Random ball query: replaces random_ball_query_cuda_kernel (.cu:58-108).  Same warp-per-query
// ordered compaction; stops as soon as nsample hits are found.
__global__ void __launch_bounds__(128) random_ball_query_kernel(
    int b, int m, int nsample, float min_radius, float max_radius, const int* __restrict__ order,
    const float* __restrict__ xyz, const float* __restrict__ new_xyz,
    const int* __restrict__ offset, const int* __restrict__ new_offset, int* __restrict__ idx,
    float* __restrict__ dist2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * 4 + warp;
    if (pt >= m) return;
    const int bt = pcm_cloud_of(pt, new_offset, b);
    const int start = bt ? __ldg(offset + bt - 1) : 0;
    const int end = __ldg(offset + bt);
    const float max_r2 = __fmul_rn(max_radius, max_radius);
    const float min_r2 = __fmul_rn(min_radius, min_radius);
    const float qx = new_xyz[(size_t)pt * 3 + 0], qy = new_xyz[(size_t)pt * 3 + 1], qz = new_xyz[(size_t)pt * 3 + 2];
    int* o_idx = idx + (size_t)pt * nsample;
    float* o_d = dist2 + (size_t)pt * nsample;
    int cnt = 0;
    for (int base = start; base < end && cnt < nsample; base += 32) {
        const int i = base + lane;
        bool hit = false;
        float d2 = 0.f;
        int src = 0;
        if (i < end) {
            src = __ldg(order + i);
            const float* p = xyz + (size_t)src * 3;
            d2 = pcm_dist2(qx - p[0], qy - p[1], qz - p[2]);
            hit = ball_hit(d2, min_r2, max_r2);
        }
        const unsigned mask = __ballot_sync(PCM_FULL_MASK, hit);
        const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < nsample) { o_d[pos] = d2; o_idx[pos] = src; }
        cnt += __popc(mask);
    }
    if (cnt > nsample) cnt = nsample;
    for (int i = cnt + lane; i < nsample; i += 32) { o_idx[i] = -1; o_d[i] = 1e10f; }
}

}  // namespace

PCM_API int pcm_knn_query(int b, int m, int nsample, const float* xyz, const float* new_xyz,
                          const int* offset, const int* new_offset, int* idx, float* dist2,
                          pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (b <= 0 || !xyz || !new_xyz || !offset || !new_offset || !idx || nsample <= 0) return PCM_EINVAL;
    if (nsample > 128) return PCM_EUNSUPPORTED;
    // PCM_KNN_V1=1: the thread-per-query kernel for every nsample (A/B timing); it also serves nsample > 31
    static const bool use_v1 = [] { const char* e = getenv("PCM_KNN_V1"); return e && e[0] == '1'; }();
    if (nsample <= 31 && !use_v1) {
        cudaStream_t st = pcm_cu_stream(stream);
        // queries per warp: two measured best at every BASELINE shape (tools/bench_pointops.py: occupancy beats sharing the
        // staged cloud between more queries); PCM_KNN_QPW = 1 | 2 | 4 | 8 overrides for timing
        static const int force = [] { const char* e = getenv("PCM_KNN_QPW"); return e ? atoi(e) : 0; }();
        switch (force) {
            case 1: return launch_knn_warp<1>(b, m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, st);
            case 4: return launch_knn_warp<4>(b, m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, st);
            case 8: return launch_knn_warp<8>(b, m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, st);
            default: return launch_knn_warp<2>(b, m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2, st);
        }
    }
    const size_t smem = (size_t)nsample * KNN_T * 8 + (size_t)KNN_TILE * 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             128 * KNN_T * 8 + KNN_TILE * 16);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    knn_kernel<<<pcm_divup(m, KNN_T), KNN_T, smem, pcm_cu_stream(stream)>>>(b, m, nsample, xyz, new_xyz, offset,
                                                                          new_offset, idx, dist2);
    return pcm_launch_status();
}

PCM_API int pcm_ball_query(int b, int m, int nsample, float min_radius, float max_radius,
                           const float* xyz, const float* new_xyz, const int* offset,
                           const int* new_offset, int* idx, float* dist2, pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (b <= 0 || !xyz || !new_xyz || !offset || !new_offset || !idx || !dist2 || nsample <= 0) return PCM_EINVAL;
    const size_t smem = (size_t)BALL_WARPS * BALL_MAX * 8;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    ball_query_kernel<<<pcm_divup(m, BALL_WARPS), BALL_WARPS * 32, smem, pcm_cu_stream(stream)>>>(
        b, m, nsample, min_radius, max_radius, xyz, new_xyz, offset, new_offset, idx, dist2);
    return pcm_launch_status();
}

PCM_API int pcm_random_ball_query(int b, int m, int nsample, float min_radius, float max_radius,
                                  const int* order, const float* xyz, const float* new_xyz,
                                  const int* offset, const int* new_offset, int* idx, float* dist2,
                                  pcm_stream_t stream) {
    if (m <= 0) return PCM_OK;
    if (b <= 0 || !order || !xyz || !new_xyz || !offset || !new_offset || !idx || !dist2 || nsample <= 0) return PCM_EINVAL;
    random_ball_query_kernel<<<pcm_divup(m, 4), 128, 0, pcm_cu_stream(stream)>>>(
        b, m, nsample, min_radius, max_radius, order, xyz, new_xyz, offset, new_offset, idx, dist2);
    return pcm_launch_status();
}
