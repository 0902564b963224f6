// ffn_fused.cu -- the transformer's feed-forward sub-block with dim_feedforward = 32 as ONE kernel each way (sm_100a).
//
//     forward :  hd = dropout(relu(x W1^T + b1)) ;  y = hd W2^T + b2                 (transformer.py:243-247,336-340)
//     backward:  dh = (dy W2) * keep_scale * [hd > 0] ;  dx = dh W1                  (dW1, dW2, db1, db2 stay with the
//                                                                                     grouped weight-gradient queue)
// The reference configuration (maniskill2_act_pcd_model.yaml:30-40) has dim_feedforward = 32 against d_model = 512: per
// token row the block moves 1 KB in (bf16 x) and 2 KB out (fp32 y) for 65 kFLOP -- 32 FLOP/B, an HBM-bound op whose
// contractions are too thin (N = 32 or K = 32) for a 128 x N x 16 tcgen05 pipeline: run as three launches (GEMM -> dropout
// -> GEMM) it costs 50 / 46 us (forward / backward) per 32 960-row layer against a 16 us traffic floor, and 28 / 26 us per
// 6 400-row layer (launch / tail latency); these kernels: 28.6 / 28.7 us and 13.0 / 13.0 us (tools/ffn_micro.py, CUDA events
// around one cold-L2 launch; ncu 25.7 / 24.2 and 11.5 / 11.0 us).  One warp owns a 16-row tile end to end: x tile -> shared
// memory (cp.async, XOR-swizzled 16-byte chunks; the next tile is prefetched into the same buffer once the first contraction
// has consumed it), both weight matrices and b2 resident in shared memory (2 x 32 KB + 2 KB, staged with cp.async), the two
// contractions on warp-level tensor-core MMAs (mma.sync.m16n8k16 bf16 -> fp32; the op needs < 1 % of the tensor peak), the
// hidden tile never leaves registers (accumulator fragments are re-packed as the A fragments of the second MMA), y / dx
// leave as full 32-byte sectors.
#include "common.cuh"

namespace {

constexpr int FF_HD = 32;        // dim_feedforward handled by this kernel
constexpr int FF_WARPS = 8;
constexpr int FF_TILE_ROWS = 16;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Shared-memory layouts (bf16).  "Wide" matrices (row = E elements = E/8 16-byte chunks): chunk c of row r is stored at
// chunk position c ^ (r & 7) -- eight consecutive rows read at one chunk column (ldmatrix) then hit eight different bank
// groups.  "Narrow" matrices (row = 32 elements = 4 chunks, 64 B): chunk c of row r at c ^ ((r >> 1) & 3).
__device__ __forceinline__ uint32_t wide_off(int r, int c, int E) { return (uint32_t)(r * E * 2 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t narrow_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// copy a row-major bf16 matrix (rows x cols, cols % 8 == 0) into shared memory with the swizzle above: asynchronous 16-byte
// copies, all in flight at once (the caller waits with cp_async_wait_all + __syncthreads)
__device__ __forceinline__ void stage_wide(const __nv_bfloat16* __restrict__ w, int rows, int cols, uint8_t* dst) {
    const int cpr = cols / 8, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = smem_addr(dst);
    for (int r = warp; r < rows; r += FF_WARPS)
        for (int c = lane; c < cpr; c += 32) cp_async16(base + wide_off(r, c, cols), w + (size_t)r * cols + c * 8, 16);
}
__device__ __forceinline__ void stage_narrow(const __nv_bfloat16* __restrict__ w, int rows, uint8_t* dst) {
    const uint32_t base = smem_addr(dst);
    for (int i = threadIdx.x; i < rows * 4; i += FF_WARPS * 32) cp_async16(base + narrow_off(i >> 2, i & 3), w + (size_t)i * 8, 16);
}

// load a 16-row tile of a (rows x E) bf16 matrix (row pitch ld) into the warp's swizzled tile; rows past `rows` are zero-filled
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ src, long ld, long row0, long rows, int E, uint8_t* tile,
                                          int lane) {
    const int cpr = E / 8;
    const uint32_t base = smem_addr(tile);
#pragma unroll 4
    for (int r = 0; r < FF_TILE_ROWS; ++r) {
        const long gr = row0 + r;
        const bool ok = gr < rows;
        const __nv_bfloat16* row = src + (ok ? gr : 0) * ld;
        for (int c = lane; c < cpr; c += 32) cp_async16(base + wide_off(r, c, E), row + c * 8, ok ? 16 : 0);
    }
}

struct FfnParams {
    const __nv_bfloat16* x;   // fwd: x (rows, E) ; bwd: dy (rows, E)
    long ldx;
    const __nv_bfloat16* w1;  // (32, E)
    const __nv_bfloat16* w2;  // (E, 32)
    const float* b1;          // (32)   fwd
    const float* b2;          // (E)    fwd
    __nv_bfloat16* hd;        // (rows, 32): fwd out / bwd in
    float* y;                 // fwd: y (rows, E) ; bwd: dx (rows, E)
    __nv_bfloat16* dh;        // bwd out (rows, 32)
    long rows;
    int E;
    float p_drop;
    const unsigned long long* seed_base;
    unsigned long long seed_offset;
};

// Work distribution: tile (cta, warp, round) = cta + grid * (warp + FF_WARPS * round) -- a short problem (the decoder's 400
// tiles) spreads over all SMs with a few warps each rather than filling 50 SMs.  Each warp prefetches its next tile (cp.async
// into the same buffer) as soon as the first contraction has consumed the current one, so the load overlaps the second
// contraction and the stores.
__global__ void __launch_bounds__(FF_WARPS * 32, 1) ffn32_fwd_kernel(const FfnParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int E = p.E;
    uint8_t* sW1 = smem;                          // 32 x E  (wide)
    uint8_t* sW2 = sW1 + (size_t)FF_HD * E * 2;   // E x 32  (narrow)
    uint8_t* sX = sW2 + (size_t)E * FF_HD * 2;    // FF_WARPS tiles of 16 x E (wide)
    float* sB2 = reinterpret_cast<float*>(sX + (size_t)FF_WARPS * FF_TILE_ROWS * E * 2);  // E
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint8_t* tile = sX + (size_t)warp * FF_TILE_ROWS * E * 2;
    const uint32_t tile_a = smem_addr(tile), w1_a = smem_addr(sW1), w2_a = smem_addr(sW2);
    const long n_tiles = (p.rows + FF_TILE_ROWS - 1) / FF_TILE_ROWS;
    const long stride = (long)gridDim.x * FF_WARPS;
    long tl = blockIdx.x + (long)gridDim.x * warp;
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    if (tl < n_tiles) load_tile(p.x, p.ldx, tl * FF_TILE_ROWS, p.rows, E, tile, lane);
    stage_wide(p.w1, FF_HD, E, sW1);
    stage_narrow(p.w2, E, sW2);
    for (int i = threadIdx.x; i < E; i += FF_WARPS * 32) cp_async4(smem_addr(sB2 + i), p.b2 + i);
    const unsigned long long seed = (p.seed_base ? *p.seed_base : 0ULL) * 0xD1342543DE82EF95ULL + p.seed_offset;
    const uint32_t thr16 = pcm_drop_thr16(p.p_drop);
    const float ks = p.p_drop > 0.f ? pcm_keep_scale(thr16) : 1.0f;
    float b1v[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { b1v[nt][0] = p.b1[nt * 8 + 2 * t]; b1v[nt][1] = p.b1[nt * 8 + 2 * t + 1]; }
    cp_async_wait_all();
    __syncthreads();

    for (; tl < n_tiles; tl += stride) {
        const long row0 = tl * FF_TILE_ROWS;
        // ---- h = x W1^T : 16 x 32, K = E ------------------------------------------------------------------
        float h[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) h[nt][0] = h[nt][1] = h[nt][2] = h[nt][3] = 0.f;
#pragma unroll 4
        for (int kc = 0; kc < E / 8; kc += 2) {  // one k-step = 16 columns = 2 chunks
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(tile_a + wide_off(lane & 15, kc + (lane >> 4), E), a0, a1, a2, a3);
#pragma unroll
            for (int np = 0; np < 2; ++np) {  // two n-tiles (16 hidden units) per ldmatrix
                uint32_t b0, b1, b2, b3;
                // matrices: (n 0-7, k lo), (n 0-7, k hi), (n 8-15, k lo), (n 8-15, k hi)
                ldmatrix_x4(w1_a + wide_off(np * 16 + (lane & 7) + ((lane >> 4) << 3), kc + ((lane >> 3) & 1), E), b0, b1, b2, b3);
                mma_bf16(h[np * 2], a0, a1, a2, a3, b0, b1);
                mma_bf16(h[np * 2 + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        __syncwarp();  // every lane has read its fragments of this tile: the buffer can take the next one
        if (tl + stride < n_tiles) load_tile(p.x, p.ldx, (tl + stride) * FF_TILE_ROWS, p.rows, E, tile, lane);
        // ---- bias + ReLU + dropout, re-pack as the A fragments of the second contraction; store hd ------------
        const long r_lo = row0 + g, r_hi = row0 + g + 8;
        uint32_t rs_lo = 0, rs_hi = 0;
        if (p.p_drop > 0.f) {
            rs_lo = pcm_row_seed(seed, (unsigned long long)r_lo * FF_HD);
            rs_hi = pcm_row_seed(seed, (unsigned long long)r_hi * FF_HD);
        }
        uint32_t hp[4][2];  // [n-tile][row half]: packed bf16 pair (cols nt*8 + 2t, +1)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float v0 = fmaxf(h[nt][0] + b1v[nt][0], 0.f), v1 = fmaxf(h[nt][1] + b1v[nt][1], 0.f);
            float v2 = fmaxf(h[nt][2] + b1v[nt][0], 0.f), v3 = fmaxf(h[nt][3] + b1v[nt][1], 0.f);
            if (p.p_drop > 0.f) {  // one hash per element pair: low / high 16 bits decide the two elements
                const uint32_t hl = pcm_pair_bits(rs_lo, nt * 4 + t), hh = pcm_pair_bits(rs_hi, nt * 4 + t);
                v0 = (hl & 0xFFFFu) >= thr16 ? v0 * ks : 0.f;
                v1 = (hl >> 16) >= thr16 ? v1 * ks : 0.f;
                v2 = (hh & 0xFFFFu) >= thr16 ? v2 * ks : 0.f;
                v3 = (hh >> 16) >= thr16 ? v3 * ks : 0.f;
            }
            hp[nt][0] = pack2(v0, v1);
            hp[nt][1] = pack2(v2, v3);
            if (r_lo < p.rows) *reinterpret_cast<uint32_t*>(p.hd + r_lo * FF_HD + nt * 8 + 2 * t) = hp[nt][0];
            if (r_hi < p.rows) *reinterpret_cast<uint32_t*>(p.hd + r_hi * FF_HD + nt * 8 + 2 * t) = hp[nt][1];
        }
        // ---- y = hd W2^T + b2 : 16 x E, K = 32, in chunks of 64 output columns (accumulators start at the bias) ----
        float* y_lo = p.y + r_lo * E + 2 * t;
        float* y_hi = p.y + r_hi * E + 2 * t;
        const bool ok_lo = r_lo < p.rows, ok_hi = r_hi < p.rows;
        for (int n0 = 0; n0 < E; n0 += 64) {
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 bb = *reinterpret_cast<const float2*>(sB2 + n0 + nt * 8 + 2 * t);
                acc[nt][0] = acc[nt][2] = bb.x;
                acc[nt][1] = acc[nt][3] = bb.y;
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                uint32_t b0, b1, b2, b3;  // (n, k 0-7), (n, k 8-15), (n, k 16-23), (n, k 24-31)
                ldmatrix_x4(w2_a + narrow_off(n0 + nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
                mma_bf16(acc[nt], hp[0][0], hp[0][1], hp[1][0], hp[1][1], b0, b1);
                mma_bf16(acc[nt], hp[2][0], hp[2][1], hp[3][0], hp[3][1], b2, b3);
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (ok_lo) *reinterpret_cast<float2*>(y_lo + n0 + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
                if (ok_hi) *reinterpret_cast<float2*>(y_hi + n0 + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
        cp_async_wait_all();
        __syncwarp();
    }
}

__global__ void __launch_bounds__(FF_WARPS * 32, 1) ffn32_bwd_kernel(const FfnParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int E = p.E;
    uint8_t* sW1 = smem;                          // 32 x E  (wide):   B of dx = dh W1, read transposed
    uint8_t* sW2 = sW1 + (size_t)FF_HD * E * 2;   // E x 32  (narrow): B of dhd = dy W2, read transposed
    uint8_t* sX = sW2 + (size_t)E * FF_HD * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint8_t* tile = sX + (size_t)warp * FF_TILE_ROWS * E * 2;
    const uint32_t tile_a = smem_addr(tile), w1_a = smem_addr(sW1), w2_a = smem_addr(sW2);
    const float ks = p.p_drop > 0.f ? pcm_keep_scale(pcm_drop_thr16(p.p_drop)) : 1.0f;
    const long n_tiles = (p.rows + FF_TILE_ROWS - 1) / FF_TILE_ROWS;
    const long stride = (long)gridDim.x * FF_WARPS;
    long tl = blockIdx.x + (long)gridDim.x * warp;
    pcm_pdl_launch_dependents();
    pcm_pdl_wait();
    if (tl < n_tiles) load_tile(p.x, p.ldx, tl * FF_TILE_ROWS, p.rows, E, tile, lane);  // dy tile
    stage_wide(p.w1, FF_HD, E, sW1);
    stage_narrow(p.w2, E, sW2);
    cp_async_wait_all();
    __syncthreads();

    for (; tl < n_tiles; tl += stride) {
        const long row0 = tl * FF_TILE_ROWS;
        const long r_lo = row0 + g, r_hi = row0 + g + 8;
        const bool ok_lo = r_lo < p.rows, ok_hi = r_hi < p.rows;
        uint32_t hdv[4][2];  // saved hidden activation (only its sign matters: dropped / clipped units are exactly 0)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            hdv[nt][0] = ok_lo ? *reinterpret_cast<const uint32_t*>(p.hd + r_lo * FF_HD + nt * 8 + 2 * t) : 0u;
            hdv[nt][1] = ok_hi ? *reinterpret_cast<const uint32_t*>(p.hd + r_hi * FF_HD + nt * 8 + 2 * t) : 0u;
        }
        // ---- dhd = dy W2 : 16 x 32, K = E; B[k][n] = W2[k][n] is row-major in smem -> transposed fragment loads ----
        float d[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
#pragma unroll 4
        for (int kc = 0; kc < E / 8; kc += 2) {
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(tile_a + wide_off(lane & 15, kc + (lane >> 4), E), a0, a1, a2, a3);
            const int k0 = kc * 8;
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t b0, b1, b2, b3;
                // matrices: (k lo 8, n-tile 2np), (k hi 8, n-tile 2np), (k lo 8, n-tile 2np+1), (k hi 8, n-tile 2np+1)
                ldmatrix_x4_trans(w2_a + narrow_off(k0 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)), b0, b1, b2, b3);
                mma_bf16(d[np * 2], a0, a1, a2, a3, b0, b1);
                mma_bf16(d[np * 2 + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        __syncwarp();  // the dy tile is consumed: prefetch the next one into the same buffer
        if (tl + stride < n_tiles) load_tile(p.x, p.ldx, (tl + stride) * FF_TILE_ROWS, p.rows, E, tile, lane);
        // ---- gate: dh = dhd * keep_scale * [hd > 0]; store dh (operand of dW1 / db1), re-pack as A fragments ----
        uint32_t dp[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const __nv_bfloat162 hl = *reinterpret_cast<const __nv_bfloat162*>(&hdv[nt][0]);
            const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&hdv[nt][1]);
            const float2 fl = __bfloat1622float2(hl), fh = __bfloat1622float2(hh);
            dp[nt][0] = pack2(fl.x > 0.f ? d[nt][0] * ks : 0.f, fl.y > 0.f ? d[nt][1] * ks : 0.f);
            dp[nt][1] = pack2(fh.x > 0.f ? d[nt][2] * ks : 0.f, fh.y > 0.f ? d[nt][3] * ks : 0.f);
            if (ok_lo) *reinterpret_cast<uint32_t*>(p.dh + r_lo * FF_HD + nt * 8 + 2 * t) = dp[nt][0];
            if (ok_hi) *reinterpret_cast<uint32_t*>(p.dh + r_hi * FF_HD + nt * 8 + 2 * t) = dp[nt][1];
        }
        // ---- dx = dh W1 : 16 x E, K = 32; B[k][n] = W1[k][n] row-major (k = hidden unit) -> transposed loads ------
        float* y_lo = p.y + r_lo * E + 2 * t;
        float* y_hi = p.y + r_hi * E + 2 * t;
        for (int n0 = 0; n0 < E; n0 += 64) {
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
                uint32_t b0, b1, b2, b3;  // (k 0-7, n), (k 8-15, n), (k 16-23, n), (k 24-31, n): rows of W1 = lane, chunk = n-tile
                ldmatrix_x4_trans(w1_a + wide_off(lane, (n0 >> 3) + nt, E), b0, b1, b2, b3);
                mma_bf16(acc[nt], dp[0][0], dp[0][1], dp[1][0], dp[1][1], b0, b1);
                mma_bf16(acc[nt], dp[2][0], dp[2][1], dp[3][0], dp[3][1], b2, b3);
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (ok_lo) *reinterpret_cast<float2*>(y_lo + n0 + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
                if (ok_hi) *reinterpret_cast<float2*>(y_hi + n0 + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
        cp_async_wait_all();
        __syncwarp();
    }
}

inline size_t ffn_smem(int E) { return (size_t)2 * FF_HD * E * 2 + (size_t)FF_WARPS * FF_TILE_ROWS * E * 2 + (size_t)E * 4; }
inline int ffn_grid(long rows) {
    const long tiles = (rows + FF_TILE_ROWS - 1) / FF_TILE_ROWS;
    return (int)(tiles < 148 ? (tiles > 0 ? tiles : 1) : 148);
}
inline int ffn_check(long long rows, int E, int Hd, const void* a, const void* b, const void* c, long long ld) {
    if (!a || !b || !c) return PCM_EINVAL;
    if (Hd != FF_HD || (E % 64) || E < 64 || E > 512 || (ld % 8)) return PCM_EUNSUPPORTED;  // E = 512: 64 KB weights + 128 KB tiles
    (void)rows;
    return PCM_OK;
}

}  // namespace

// Fused FFN forward for dim_feedforward = 32: x (rows, E) bf16 (row pitch ldx), W1 (32, E) bf16, b1 (32) fp32, W2 (E, 32) bf16,
// b2 (E) fp32 -> y (rows, E) fp32 and the dropped hidden activation hd (rows, 32) bf16 (saved for the backward; units that
// were clipped by the ReLU or dropped are exactly 0).  E % 64 == 0, 64 <= E <= 512.  p_drop = 0: no dropout.
PCM_API int pcm_ffn32_fwd(long long rows, int E, int Hd, const void* x, long long ldx, const void* w1, const float* b1, const void* w2,
                          const float* b2, float p_drop, const unsigned long long* seed_base, unsigned long long seed_offset,
                          void* hd, float* y, pcm_stream_t stream) {
    if (rows <= 0) return PCM_OK;
    int r = ffn_check(rows, E, Hd, x, w1, w2, ldx);
    if (r) return r;
    if (!b1 || !b2 || !hd || !y || p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    FfnParams p{reinterpret_cast<const __nv_bfloat16*>(x), (long)ldx, reinterpret_cast<const __nv_bfloat16*>(w1),
                reinterpret_cast<const __nv_bfloat16*>(w2), b1, b2, reinterpret_cast<__nv_bfloat16*>(hd), y, nullptr, (long)rows, E,
                p_drop, seed_base, seed_offset};
    const size_t smem = ffn_smem(E);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ffn32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffn_smem(512));
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    cudaError_t e = pcm_launch(ffn32_fwd_kernel, dim3(ffn_grid(rows)), dim3(FF_WARPS * 32), smem, pcm_cu_stream(stream), p);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}

// Fused FFN backward (activation gradients): dy (rows, E) bf16 (pitch lddy), hd (rows, 32) bf16 from the forward ->
// dh (rows, 32) bf16 = (dy W2) * keep_scale(p_drop) * [hd > 0] and dx (rows, E) fp32 = dh W1.
PCM_API int pcm_ffn32_bwd(long long rows, int E, int Hd, const void* dy, long long lddy, const void* hd, const void* w1,
                          const void* w2, float p_drop, void* dh, float* dx, pcm_stream_t stream) {
    if (rows <= 0) return PCM_OK;
    int r = ffn_check(rows, E, Hd, dy, w1, w2, lddy);
    if (r) return r;
    if (!hd || !dh || !dx || p_drop < 0.f || p_drop >= 1.f) return PCM_EINVAL;
    FfnParams p{reinterpret_cast<const __nv_bfloat16*>(dy), (long)lddy, reinterpret_cast<const __nv_bfloat16*>(w1),
                reinterpret_cast<const __nv_bfloat16*>(w2), nullptr, nullptr,
                const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(hd)), dx, reinterpret_cast<__nv_bfloat16*>(dh),
                (long)rows, E, p_drop, nullptr, 0ULL};
    const size_t smem = ffn_smem(E);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ffn32_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffn_smem(512));
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    cudaError_t e = pcm_launch(ffn32_bwd_kernel, dim3(ffn_grid(rows)), dim3(FF_WARPS * 32), smem, pcm_cu_stream(stream), p);
    if (e != cudaSuccess) return (int)e;
    return pcm_launch_status();
}
