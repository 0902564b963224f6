// common.cuh -- shared helpers for the sm_100a kernels of libpcm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <utility>
#include "../../include/pcm_b200.h"

#define PCM_API extern "C" __attribute__((visibility("default")))

#define PCM_FULL_MASK 0xffffffffu

// every kernel launch in the library is followed by pcm_launch_status(): it bumps the process-wide
// launch counter (pcm_launch_count(), bench.py's `gpu_launches`) and returns the launch status.
extern long long g_pcm_launch_count;
static inline int pcm_launch_status() { ++g_pcm_launch_count; return (int)cudaPeekAtLastError(); }
static inline cudaStream_t pcm_cu_stream(pcm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int pcm_divup(long a, long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The training step is ~1100 short kernels (average 14 us) replayed from a CUDA graph: launch
// latency and per-kernel prologues (tensor-map fetch, barrier init, TMEM allocation) are a
// double-digit share of it.  Kernels launched through pcm_launch() carry the
// programmatic-stream-serialization attribute: their CTAs may become resident and run their
// prologue while the preceding kernel is still draining; pcm_pdl_wait() -- executed before the first
// global-memory access -- blocks until every prerequisite grid has completed and flushed, so memory
// semantics are exactly those of ordinary stream order.  pcm_pdl_launch_dependents() lets the NEXT
// kernel start its own prologue early.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pcm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pcm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pcm_pdl_enabled();  // PCM_PDL=1 in the environment switches the attribute on (A/B timing; measured neutral)

template <typename... KArgs, typename... Args>
static inline cudaError_t pcm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pcm_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Reference block-size rule (libs/pointops/src/cuda_utils.h:11-14): largest power of two
// <= work_size, capped at 1024.  Host-side, same libm expression as the reference launcher.
int pcm_ref_opt_n_threads(int work_size);

// Squared distance with the reference's FMA contraction: nvcc turns its
// `dx*dx + dy*dy + dz*dz` into FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.) (read off the reference's
// sm_100a SASS and confirmed bit-exact against its kNN distances on a B200).  Written with
// explicit intrinsics so that neither -fmad nor instruction scheduling can change the rounding.
__device__ __forceinline__ float pcm_dist2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// cloud id of element `i` given cumulative end offsets (first c with i < offset[c]);
// equals the reference's linear get_bt_idx (knn_query_cuda_kernel.cu:45-56) for i < offset[b-1].
__device__ __forceinline__ int pcm_cloud_of(int i, const int* __restrict__ offset, int b) {
    int lo = 0, hi = b - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (i < __ldg(offset + mid)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): two IEEE fp32 operations per instruction, bit-identical to the scalar
// forms.  Used where a kernel is bound by its instruction count (set-abstraction gather, attention softmax).
__device__ __forceinline__ float2 pcm_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 pcm_fadd2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 pcm_fmul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}

// ---- counter-based dropout RNG shared by the attention / LayerNorm kernels --------------------
// A 64-bit splitmix mix per ROW (seed, row base index) and one cheap 32-bit hash ("lowbias32") per
// PAIR of elements: low / high 16 bits decide the two elements.  The drop probability actually
// realised is thr16 / 65536; kernels scale survivors by 65536 / (65536 - thr16), so the estimator
// is exactly unbiased.  Stateless: the backward kernels regenerate the identical mask.
__device__ __forceinline__ uint32_t pcm_row_seed(unsigned long long seed, unsigned long long row_base) {
    unsigned long long x = seed + row_base * 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return (uint32_t)(x >> 32) ^ (uint32_t)x;
}
__device__ __forceinline__ uint32_t pcm_pair_bits(uint32_t row_seed, uint32_t pair) {
    uint32_t x = row_seed ^ (pair * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
// Attention dropout (csrc/flash_attn.cu): 8 consecutive keys of a row share ONE strong hash
// (group g = key >> 3); key k of the group uses that hash advanced k times by a 32-bit LCG, and
// is kept iff the top 16 bits are >= thr16, i.e. x >= thr16 << 16 as one unsigned compare.
// 2 integer ops per element + 10 per group instead of 5 + 1.5 for one hash per pair.
#define PCM_LCG_A 0x2C9277B5u
#define PCM_LCG_C 0x9E3779B9u
__device__ __forceinline__ uint32_t pcm_lcg_next(uint32_t x) { return x * PCM_LCG_A + PCM_LCG_C; }
__host__ __device__ __forceinline__ uint32_t pcm_drop_thr16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
__host__ __device__ __forceinline__ float pcm_keep_scale(uint32_t thr16) { return 65536.0f / (65536.0f - (float)thr16); }
