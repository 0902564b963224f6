// fps.cu -- farthest point sampling for sm_100a.
//
// Replaces farthest_point_sampling_cuda_kernel (reference libs/pointops/src/sampling/
// sampling_cuda_kernel.cu:14-129).  Same result bit for bit, different machine mapping:
//   * one CTA per cloud; every thread keeps its points' xyz AND running min-distance in
//     REGISTERS (the reference re-reads xyz and read-modify-writes `tmp` in global memory every
//     round), so a round touches no memory except one L1-resident 12-byte broadcast load;
//   * the block-wide arg-max is two REDUX (warp reduce) instructions per level and ONE
//     __syncthreads per round (the reference does a 10-level shared-memory tree with 11 barriers);
//   * ties are resolved by an explicit priority key that reproduces the reference's implicit
//     rule: its thread `t` of a BS-wide block scans points t, t+BS, ... keeping the first strict
//     maximum, and its shared-memory tree (fold upper half onto lower half, lower slot wins a tie)
//     lets the thread with the smallest BIT-REVERSED id survive, so among equal distances the
//     winner minimises (bitrev(rel % BS), rel / BS) with rel = k - start_n and
//     BS = opt_n_threads(n_max) (cuda_utils.h:11-14).
// FPS is a chain of M-1 dependent block-wide reductions: it is latency-bound, never HBM-bound
// (algorithmic traffic 12N+4M bytes per cloud).
#include "common.cuh"
#include <cstdlib>
#include <math.h>

long long g_pcm_launch_count = 0;
PCM_API long long pcm_launch_count(void) { return g_pcm_launch_count; }

bool pcm_pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("PCM_PDL");
        on = (e && e[0] == '1') ? 1 : 0;  // measured neutral inside the CUDA-graph replay of the step: opt-in
    }
    return on == 1;
}

int pcm_ref_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

namespace {

constexpr int kKeyQBits = 16;  // key = bitrev(rel % BS) << 16 | (rel / BS)
constexpr int kNegOneBits = (int)0xBF800000;  // __float_as_int(-1.0f): below every d >= +0 as a signed int

__device__ __forceinline__ unsigned fps_brev(unsigned r, int bs_log2) {
    return bs_log2 ? (__brev(r) >> (32 - bs_log2)) : 0u;
}
__device__ __forceinline__ unsigned fps_key(int rel, int bs_log2) {
    const unsigned r = (unsigned)rel & ((1u << bs_log2) - 1u);
    return (fps_brev(r, bs_log2) << kKeyQBits) | (unsigned)(rel >> bs_log2);
}
__device__ __forceinline__ int fps_rel_of_key(unsigned key, int bs_log2) {
    const unsigned r = fps_brev(key >> kKeyQBits, bs_log2);
    return (int)(((key & ((1u << kKeyQBits) - 1u)) << bs_log2) | r);
}

// Block-wide (max distance, then min key) with a single barrier.  Every warp ends up holding
// the result, so no second barrier / broadcast is needed.  `buf` alternates per round.
template <int NW>
__device__ __forceinline__ unsigned fps_block_argmax(float best, unsigned bkey, int (*s_d)[32],
                                                     unsigned (*s_k)[32], int buf, int lane, int warp) {
    const int bits = __float_as_int(best);
    const int wmax = __reduce_max_sync(PCM_FULL_MASK, bits);
    const unsigned wkey = __reduce_min_sync(PCM_FULL_MASK, bits == wmax ? bkey : 0xFFFFFFFFu);
    if (NW == 1) return wkey;
    if (lane == 0) { s_d[buf][warp] = wmax; s_k[buf][warp] = wkey; }
    __syncthreads();
    const int vd = lane < NW ? s_d[buf][lane] : kNegOneBits;
    const unsigned vk = lane < NW ? s_k[buf][lane] : 0xFFFFFFFFu;
    const int bmax = __reduce_max_sync(PCM_FULL_MASK, vd);
    return __reduce_min_sync(PCM_FULL_MASK, vd == bmax ? vk : 0xFFFFFFFFu);
}

// Register-resident variant: cloud size <= T * PPT.
template <int T, int PPT>
__global__ void __launch_bounds__(T) fps_regs_kernel(const float* __restrict__ xyz,
                                                     const int* __restrict__ offset,
                                                     const int* __restrict__ new_offset,
                                                     int bs_log2, int* __restrict__ idx) {
    constexpr int NW = T / 32;
    __shared__ int s_d[2][32];
    __shared__ unsigned s_k[2][32];
    const int bid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start_n = bid ? offset[bid - 1] : 0;
    const int end_n = offset[bid];
    const int start_m = bid ? new_offset[bid - 1] : 0;
    const int end_m = new_offset[bid];
    int n = end_n - start_n;
    if (end_m <= start_m) return;
    if (tid == 0) idx[start_m] = start_n;
    if (n <= 0) {  // reference: every thread reports (-1, start_n) -> start_n wins each round
        for (int j = start_m + 1 + tid; j < end_m; j += T) idx[j] = start_n;
        return;
    }
    if (n > T * PPT) n = T * PPT;  // launcher guarantees this never triggers; memory-safety clamp

    float px[PPT], py[PPT], pz[PPT], md[PPT];
    unsigned key[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int rel = tid + p * T;
        const bool valid = rel < n;
        const float* q = xyz + (size_t)(start_n + (valid ? rel : 0)) * 3;
        px[p] = q[0]; py[p] = q[1]; pz[p] = q[2];
        md[p] = valid ? 1e10f : -1.0f;  // invalid slots stay at -1 forever (min(d,-1) = -1)
        key[p] = fps_key(rel, bs_log2);
    }
    float ox = xyz[(size_t)start_n * 3 + 0], oy = xyz[(size_t)start_n * 3 + 1], oz = xyz[(size_t)start_n * 3 + 2];

    for (int j = start_m + 1; j < end_m; ++j) {
        float best = -1.0f;
        unsigned bkey = 0xFFFFFFFFu;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const float d = pcm_dist2(px[p] - ox, py[p] - oy, pz[p] - oz);
            const float d2 = fminf(d, md[p]);
            md[p] = d2;
            const bool better = (d2 > best) || (d2 == best && key[p] < bkey);
            best = better ? d2 : best;
            bkey = better ? key[p] : bkey;
        }
        const unsigned wkey = fps_block_argmax<NW>(best, bkey, s_d, s_k, j & 1, lane, warp);
        const int old = start_n + fps_rel_of_key(wkey, bs_log2);
        const float* q = xyz + (size_t)old * 3;
        ox = __ldg(q + 0); oy = __ldg(q + 1); oz = __ldg(q + 2);
        if (tid == 0) idx[j] = old;
    }
}

// Generic variant for clouds larger than 8192 points: running minima in the caller's `tmp`
// buffer (pre-filled with 1e10, reference convention), still one barrier per round.
template <int T>
__global__ void __launch_bounds__(T) fps_generic_kernel(const float* __restrict__ xyz,
                                                        const int* __restrict__ offset,
                                                        const int* __restrict__ new_offset,
                                                        int bs_log2, float* __restrict__ tmp,
                                                        int* __restrict__ idx) {
    constexpr int NW = T / 32;
    __shared__ int s_d[2][32];
    __shared__ unsigned s_k[2][32];
    const int bid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start_n = bid ? offset[bid - 1] : 0;
    const int end_n = offset[bid];
    const int start_m = bid ? new_offset[bid - 1] : 0;
    const int end_m = new_offset[bid];
    const int n = end_n - start_n;
    if (end_m <= start_m) return;
    if (tid == 0) idx[start_m] = start_n;
    if (n <= 0) {
        for (int j = start_m + 1 + tid; j < end_m; j += T) idx[j] = start_n;
        return;
    }
    float ox = xyz[(size_t)start_n * 3 + 0], oy = xyz[(size_t)start_n * 3 + 1], oz = xyz[(size_t)start_n * 3 + 2];
    for (int j = start_m + 1; j < end_m; ++j) {
        float best = -1.0f;
        unsigned bkey = 0xFFFFFFFFu;
        for (int rel = tid; rel < n; rel += T) {
            const float* q = xyz + (size_t)(start_n + rel) * 3;
            const float d = pcm_dist2(q[0] - ox, q[1] - oy, q[2] - oz);
            const float d2 = fminf(d, tmp[start_n + rel]);
            tmp[start_n + rel] = d2;
            const unsigned k = fps_key(rel, bs_log2);
            const bool better = (d2 > best) || (d2 == best && k < bkey);
            best = better ? d2 : best;
            bkey = better ? k : bkey;
        }
        const unsigned wkey = fps_block_argmax<NW>(best, bkey, s_d, s_k, j & 1, lane, warp);
        const int old = start_n + fps_rel_of_key(wkey, bs_log2);
        const float* q = xyz + (size_t)old * 3;
        ox = __ldg(q + 0); oy = __ldg(q + 1); oz = __ldg(q + 2);
        if (tid == 0) idx[j] = old;
    }
}

int g_fps_threads_override = 0;  // tuning hook (pcm_tune_fps_threads)

template <int T>
int launch_regs(int ppt, int b, const float* xyz, const int* offset, const int* new_offset,
                int bs_log2, int* idx, cudaStream_t st) {
    switch (ppt) {
        case 1: fps_regs_kernel<T, 1><<<b, T, 0, st>>>(xyz, offset, new_offset, bs_log2, idx); break;
        case 2: fps_regs_kernel<T, 2><<<b, T, 0, st>>>(xyz, offset, new_offset, bs_log2, idx); break;
        case 3: case 4: fps_regs_kernel<T, 4><<<b, T, 0, st>>>(xyz, offset, new_offset, bs_log2, idx); break;
        default: fps_regs_kernel<T, 8><<<b, T, 0, st>>>(xyz, offset, new_offset, bs_log2, idx); break;
    }
    return pcm_launch_status();
}

}  // namespace

// Tuning hook, not part of the reference surface: force the CTA width of the register-resident
// FPS kernel (0 = automatic).  Used by bench / profiling sweeps.
PCM_API int pcm_tune_fps_threads(int threads) {
    if (threads != 0 && threads != 128 && threads != 256 && threads != 512 && threads != 1024) return PCM_EINVAL;
    g_fps_threads_override = threads;
    return PCM_OK;
}

PCM_API int pcm_farthest_point_sampling(int b, int n, const float* xyz, const int* offset,
                                        const int* new_offset, float* tmp, int* idx,
                                        pcm_stream_t stream) {
    if (b <= 0) return PCM_OK;
    if (!xyz || !offset || !new_offset || !idx || n <= 0) return PCM_EINVAL;
    cudaStream_t st = pcm_cu_stream(stream);
    const int bs = pcm_ref_opt_n_threads(n);
    int bs_log2 = 0;
    while ((1 << bs_log2) < bs) ++bs_log2;
    if (n > 8192) {
        if (!tmp) return PCM_EINVAL;
        fps_generic_kernel<1024><<<b, 1024, 0, st>>>(xyz, offset, new_offset, bs_log2, tmp, idx);
        return pcm_launch_status();
    }
    int T = g_fps_threads_override;
    if (T == 0) T = n <= 4096 ? 512 : 1024;  // measured on B200: 512-wide CTAs win for N = 1024..4096
    while (T < 1024 && (n + T - 1) / T > 8) T *= 2;
    const int ppt = (n + T - 1) / T;
    switch (T) {
        case 128: return launch_regs<128>(ppt, b, xyz, offset, new_offset, bs_log2, idx, st);
        case 256: return launch_regs<256>(ppt, b, xyz, offset, new_offset, bs_log2, idx, st);
        case 512: return launch_regs<512>(ppt, b, xyz, offset, new_offset, bs_log2, idx, st);
        default: return launch_regs<1024>(ppt, b, xyz, offset, new_offset, bs_log2, idx, st);
    }
}
