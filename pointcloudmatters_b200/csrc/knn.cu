// knn.cu -- kNN / ball / random-ball query for sm_100a.
//
// kNN replaces knn_query_cuda_kernel (reference libs/pointops/src/knn_query/
// knn_query_cuda_kernel.cu:60-104).  The reference's output ORDER under equal distances is an
// artefact of its sequential binary max-heap (strict `d2 < root` replacement, `reheap`, then
// `heap_sort`), so the only way to be bit-exact on every input is to replay that heap.  We keep
// one query per thread like the reference, but
//   * the heap lives in SHARED memory laid out [slot][thread] (bank = thread id: conflict-free
//     for any per-thread slot), not in a 1 KB/thread local-memory stack frame;
//   * the cloud is staged through shared memory in float4 tiles with coalesced loads, so the
//     inner loop is one broadcast LDS.128 + 6 FP ops per candidate instead of three global loads;
//   * the heap root is cached in a register, so the common reject path touches no memory;
//   * results leave through a block-wide transposed, fully coalesced store.
// Algorithmic traffic: 12N + 12M + 8*M*k bytes per cloud.
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
