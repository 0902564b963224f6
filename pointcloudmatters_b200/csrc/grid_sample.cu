// grid_sample.cu -- voxel-grid subsampling of a packed batch of raw clouds on the GPU.
//
// Replaces, for a whole batch at once, the per-sample CPU transform GridSamplePCD (reference
// src/data/components/transformpcd.py:684-793, hash_type "fnv") that 16 loader workers run today:
//     grid = floor(coord / grid_size) (float64 arithmetic, like numpy) ; grid -= grid.min(0) (per cloud)
//     key  = FNV64-1A over the three grid coordinates                       (:775-793)
//     one point per distinct key, output ORDERED BY ASCENDING KEY            (argsort + unique, :693-704)
// Which member of a voxel survives: the one with the smallest 32-bit priority (ties: smallest index).  priority = point
// index reproduces the reference's test-mode part 0 under a stable argsort; priority = a per-point random number is its
// train mode (uniform member per voxel; numpy's RNG stream itself is not reproducible on a GPU).
// Integer / byte work, HBM- and atomics-bound: no tensor cores.  Three kernels:
//   1. grid_voxelize: per-point grid coordinates (int32 x 3) + per-cloud minimum (atomicMin);
//   2. grid_insert:   key -> open-addressing hash table region of the point's cloud (2 x cloud size slots),
//                     atomicCAS on the 64-bit key, atomicMin on (priority << 32 | local index);
//   3. grid_compact_sort: one CTA per cloud gathers the occupied slots, sorts them by key (bitonic network in shared
//                     memory up to 8192 voxels, in global scratch beyond) and writes the selected row indices, the
//                     min-subtracted grid coordinates and the voxel count.
#include "common.cuh"

namespace {

constexpr unsigned long long FNV_OFFSET = 14695981039346656037ULL;
constexpr unsigned long long FNV_PRIME = 1099511628211ULL;
constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFULL;
constexpr int SORT_CAP = 8192;  // voxels per cloud sorted in shared memory (8192 x 12 B = 96 KB)

__device__ __forceinline__ unsigned long long fnv3(long long gx, long long gy, long long gz) {
    unsigned long long h = FNV_OFFSET;
    h *= FNV_PRIME; h ^= (unsigned long long)gx;
    h *= FNV_PRIME; h ^= (unsigned long long)gy;
    h *= FNV_PRIME; h ^= (unsigned long long)gz;
    return h;
}

__global__ void __launch_bounds__(256) grid_voxelize_kernel(const float* __restrict__ coord, const long long* __restrict__ offset, int b,
                                                            long n, double gsx, double gsy, double gsz, int f32_div,
                                                            int* __restrict__ grid, int* __restrict__ gmin) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int lo = 0, hi = b - 1;  // cloud of point i: first c with i < offset[c]
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (i < offset[mid]) hi = mid; else lo = mid + 1;
        }
        // numpy >= 2 promotes float32 array / float64 0-d array to float64; numpy 1.x (value-based casting) stays float32
        int gx, gy, gz;
        if (f32_div) {
            gx = (int)floorf(__fdiv_rn(coord[i * 3 + 0], (float)gsx));
            gy = (int)floorf(__fdiv_rn(coord[i * 3 + 1], (float)gsy));
            gz = (int)floorf(__fdiv_rn(coord[i * 3 + 2], (float)gsz));
        } else {
            gx = (int)floor((double)coord[i * 3 + 0] / gsx);
            gy = (int)floor((double)coord[i * 3 + 1] / gsy);
            gz = (int)floor((double)coord[i * 3 + 2] / gsz);
        }
        grid[i * 3 + 0] = gx; grid[i * 3 + 1] = gy; grid[i * 3 + 2] = gz;
        atomicMin(gmin + lo * 3 + 0, gx);
        atomicMin(gmin + lo * 3 + 1, gy);
        atomicMin(gmin + lo * 3 + 2, gz);
    }
}

__global__ void __launch_bounds__(256) grid_insert_kernel(const int* __restrict__ grid, const int* __restrict__ gmin,
                                                          const long long* __restrict__ offset, int b, long n,
                                                          const unsigned int* __restrict__ prio, unsigned long long* __restrict__ tkey,
                                                          unsigned long long* __restrict__ tbest) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int lo = 0, hi = b - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (i < offset[mid]) hi = mid; else lo = mid + 1;
        }
        const long start = lo > 0 ? offset[lo - 1] : 0;
        const long cap = 2 * (offset[lo] - start);  // slots of this cloud's table region [2 * start, 2 * start + cap)
        const unsigned long long key = fnv3(grid[i * 3 + 0] - gmin[lo * 3 + 0], grid[i * 3 + 1] - gmin[lo * 3 + 1],
                                            grid[i * 3 + 2] - gmin[lo * 3 + 2]);
        const unsigned int pr = prio ? prio[i] : (unsigned int)(i - start);
        const unsigned long long cand = ((unsigned long long)pr << 32) | (unsigned int)(i - start);
        unsigned long long* k = tkey + 2 * start;
        unsigned long long* v = tbest + 2 * start;
        long s = (long)((key * 0x9E3779B97F4A7C15ULL) >> 33) % cap;
        while (true) {
            const unsigned long long prev = atomicCAS(k + s, EMPTY_KEY, key);
            if (prev == EMPTY_KEY || prev == key) { atomicMin(v + s, cand); break; }
            if (++s == cap) s = 0;
        }
    }
}

__device__ __forceinline__ void bitonic_sort(unsigned long long* key, unsigned int* val, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = key[i], c = key[ixj];
                    if ((a > c) == up) {
                        key[i] = c; key[ixj] = a;
                        const unsigned int t = val[i]; val[i] = val[ixj]; val[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// one CTA per cloud.  out_idx / out_grid are written at the cloud's RAW offsets (upper-bound layout); counts[c] = voxels.
__global__ void __launch_bounds__(1024) grid_compact_sort_kernel(const unsigned long long* __restrict__ tkey,
                                                                 const unsigned long long* __restrict__ tbest,
                                                                 const int* __restrict__ grid, const int* __restrict__ gmin,
                                                                 const long long* __restrict__ offset,
                                                                 unsigned long long* __restrict__ scratch_key,
                                                                 unsigned int* __restrict__ scratch_val,
                                                                 long long* __restrict__ out_idx, long long* __restrict__ out_grid,
                                                                 int* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_count;
    const int c = blockIdx.x;
    const long start = c > 0 ? offset[c - 1] : 0;
    const long n_c = offset[c] - start;
    const long cap = 2 * n_c;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // pass 1: count occupied slots
    int local = 0;
    for (long s = threadIdx.x; s < cap; s += blockDim.x) local += tkey[2 * start + s] != EMPTY_KEY;
    atomicAdd(&s_count, local);
    __syncthreads();
    const int U = s_count;
    int pow2 = 1;
    while (pow2 < U) pow2 <<= 1;
    const bool in_smem = pow2 <= SORT_CAP;
    unsigned long long* key = in_smem ? reinterpret_cast<unsigned long long*>(smem_raw) : scratch_key + 2 * start;
    unsigned int* val = in_smem ? reinterpret_cast<unsigned int*>(smem_raw + (size_t)SORT_CAP * 8) : scratch_val + 2 * start;
    __syncthreads();
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // pass 2: gather (order irrelevant, the sort fixes it) ; pad to a power of two with +inf keys
    for (long s = threadIdx.x; s < cap; s += blockDim.x) {
        const unsigned long long k = tkey[2 * start + s];
        if (k != EMPTY_KEY) {
            const int slot = atomicAdd(&s_count, 1);
            key[slot] = k;
            val[slot] = (unsigned int)(tbest[2 * start + s] & 0xFFFFFFFFULL);  // local index of the surviving member
        }
    }
    for (int i = U + threadIdx.x; i < pow2; i += blockDim.x) { key[i] = EMPTY_KEY; val[i] = 0xFFFFFFFFu; }
    __syncthreads();
    bitonic_sort(key, val, pow2);
    for (int i = threadIdx.x; i < U; i += blockDim.x) {
        const long row = start + val[i];
        out_idx[start + i] = row;
        out_grid[(start + i) * 3 + 0] = grid[row * 3 + 0] - gmin[c * 3 + 0];
        out_grid[(start + i) * 3 + 1] = grid[row * 3 + 1] - gmin[c * 3 + 1];
        out_grid[(start + i) * 3 + 2] = grid[row * 3 + 2] - gmin[c * 3 + 2];
    }
    if (threadIdx.x == 0) counts[c] = U;
}

// dense packing after the host knows the voxel counts: row j of cloud c (new offsets) <- raw slot start_c + j
__global__ void __launch_bounds__(256) grid_gather_kernel(const long long* __restrict__ raw_offset, const long long* __restrict__ new_offset,
                                                          int b, long m, const long long* __restrict__ idx_raw,
                                                          const long long* __restrict__ grid_raw, const float* __restrict__ coord,
                                                          const float* __restrict__ feat, int fc, float feat_scale, float feat_shift,
                                                          int append_coord, float* __restrict__ coord_out,
                                                          long long* __restrict__ grid_out, float* __restrict__ feat_out,
                                                          long long* __restrict__ index_out) {
    const int oc = fc + (append_coord ? 3 : 0);
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long)gridDim.x * blockDim.x) {
        int lo = 0, hi = b - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (j < new_offset[mid]) hi = mid; else lo = mid + 1;
        }
        const long nstart = lo > 0 ? new_offset[lo - 1] : 0, rstart = lo > 0 ? raw_offset[lo - 1] : 0;
        const long slot = rstart + (j - nstart);
        const long row = idx_raw[slot];
        const float x = coord[row * 3 + 0], y = coord[row * 3 + 1], z = coord[row * 3 + 2];
        coord_out[j * 3 + 0] = x; coord_out[j * 3 + 1] = y; coord_out[j * 3 + 2] = z;
        grid_out[j * 3 + 0] = grid_raw[slot * 3 + 0]; grid_out[j * 3 + 1] = grid_raw[slot * 3 + 1]; grid_out[j * 3 + 2] = grid_raw[slot * 3 + 2];
        for (int k = 0; k < fc; ++k) feat_out[j * oc + k] = feat[row * fc + k] / feat_scale - feat_shift;
        if (append_coord) { feat_out[j * oc + fc] = x; feat_out[j * oc + fc + 1] = y; feat_out[j * oc + fc + 2] = z; }
        if (index_out) index_out[j] = row;
    }
}

inline int gs_grid(long n) { const long g = (n + 255) / 256; return (int)(g < 148L * 8 ? (g > 0 ? g : 1) : 148L * 8); }

}  // namespace

// Stage 1: voxelise + hash + per-cloud sort.  Workspace (caller-allocated): grid (n, 3) int32; gmin (b, 3) int32
// preset to INT_MAX; tkey / tbest (2n) uint64 preset to all-ones; scratch_key (2n) uint64 / scratch_val (2n) uint32 (only
// touched for clouds with more than 8192 voxels).  Outputs at RAW offsets: idx_raw (n) int64 global row of the survivor
// of the j-th voxel (ascending key) of each cloud, grid_raw (n, 3) int64 min-subtracted grid coordinates, counts (b).
PCM_API int pcm_grid_sample_select(int b, long long n, const float* coord, const long long* offset, double gsx, double gsy,
                                   double gsz, int f32_div, const unsigned int* prio, int* grid, int* gmin, unsigned long long* tkey,
                                   unsigned long long* tbest, unsigned long long* scratch_key, unsigned int* scratch_val,
                                   long long* idx_raw, long long* grid_raw, int* counts, pcm_stream_t stream) {
    if (b <= 0 || n <= 0) return PCM_OK;
    if (!coord || !offset || !grid || !gmin || !tkey || !tbest || !scratch_key || !scratch_val || !idx_raw || !grid_raw || !counts)
        return PCM_EINVAL;
    if (!(gsx > 0) || !(gsy > 0) || !(gsz > 0)) return PCM_EINVAL;
    cudaStream_t st = pcm_cu_stream(stream);
    grid_voxelize_kernel<<<gs_grid(n), 256, 0, st>>>(coord, offset, b, (long)n, gsx, gsy, gsz, f32_div, grid, gmin);
    int r = pcm_launch_status();
    if (r) return r;
    grid_insert_kernel<<<gs_grid(n), 256, 0, st>>>(grid, gmin, offset, b, (long)n, prio, tkey, tbest);
    r = pcm_launch_status();
    if (r) return r;
    const size_t smem = (size_t)SORT_CAP * 12;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(grid_compact_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    grid_compact_sort_kernel<<<b, 1024, smem, st>>>(tkey, tbest, grid, gmin, offset, scratch_key, scratch_val, idx_raw, grid_raw, counts);
    return pcm_launch_status();
}

// Stage 2: dense packing.  feat (n, fc) raw per-point features (e.g. uint8-valued colours as float); feat_out[:, :fc] =
// feat / feat_scale - feat_shift (NormalizeColorPCD: / 127.5 - 1, transformpcd.py), followed by the coordinates when
// append_coord (CollectPCD feat_keys [color, coord]).  index_out (m) optional: the selected raw rows.
PCM_API int pcm_grid_sample_gather(int b, long long m, const long long* raw_offset, const long long* new_offset,
                                   const long long* idx_raw, const long long* grid_raw, const float* coord, const float* feat,
                                   int fc, float feat_scale, float feat_shift, int append_coord, float* coord_out,
                                   long long* grid_out, float* feat_out, long long* index_out, pcm_stream_t stream) {
    if (b <= 0 || m <= 0) return PCM_OK;
    if (!raw_offset || !new_offset || !idx_raw || !grid_raw || !coord || !coord_out || !grid_out || (fc > 0 && (!feat || !feat_out)))
        return PCM_EINVAL;
    if (feat_scale == 0.f) return PCM_EINVAL;
    grid_gather_kernel<<<gs_grid(m), 256, 0, pcm_cu_stream(stream)>>>(raw_offset, new_offset, b, (long)m, idx_raw, grid_raw, coord, feat,
                                                                     fc, feat_scale, feat_shift, append_coord, coord_out, grid_out,
                                                                     feat_out, index_out);
    return pcm_launch_status();
}
