// tokens.cu -- observation-token assembly and the action heads + loss of the ACT policy for sm_100a.
//
// (1) pcm_coord_embed_sine_tokens: ACTPCD.coord_embedding_sine (reference src/models/components/act/act.py:467-506,
//     normalize=False) evaluated on the sampled coordinates and written STRAIGHT into the transformer's seq-first
//     positional tensor (S, B, E) = [additional_pos_embed rows broadcast over the batch ; sine rows]
//     (transformer.py:75-88 builds it with flatten / permute / repeat / cat passes over 67 MB at cfg-2).
// (2) pcm_fill_head_rows: the latent / proprio / goal rows of the token tensor (transformer.py:89-92) plus their bf16
//     operand copies; the point rows are written by pcm_sa_output_tokens (sa_fused.cu).
// (3) pcm_act_heads_loss_fwd / _bwd: action_head + is_pad_head + masked MSE + KL (act.py:255-291; RLBench variant
//     :770-825 with sigmoid gripper / collision outputs and a weighted position loss; loss/misc.py:11-26) as one kernel
//     each way.  All three are HBM / latency-bound elementwise work: 128-bit accesses, one warp per token row.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    return pk;
}

// one CTA per token row r = s * B + b; thread t owns channels 4t .. 4t+3 (E <= 4 * blockDim)
__global__ void __launch_bounds__(256) coord_embed_sine_tokens_kernel(const float* __restrict__ coord, const float* __restrict__ dim_t,
                                                                      const float* __restrict__ add_pos, int per, int batch, int head,
                                                                      int E, int npf, float* __restrict__ pos) {
    const long r = blockIdx.x;
    const int s = (int)(r / batch), b = (int)(r % batch);
    float* dst = pos + r * E;
    if (s < head) {
        for (int c = threadIdx.x * 4; c < E; c += blockDim.x * 4)
            *reinterpret_cast<float4*>(dst + c) = *reinterpret_cast<const float4*>(add_pos + (size_t)s * E + c);
        return;
    }
    const float* xyz = coord + ((size_t)b * per + (s - head)) * 3;
    const float x = __ldg(xyz), y = __ldg(xyz + 1), z = __ldg(xyz + 2);
    for (int c = threadIdx.x * 4; c < E; c += blockDim.x * 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = c + u;
            const int axis = j / npf, jj = j - axis * npf;
            if (axis >= 3) { v[u] = 0.f; continue; }  // E - 3 * npf zero pad channels (act.py:505)
            // per axis the reference lays out [sin of the even dim_t entries | cos of the odd ones]
            // (torch.stack(..., dim=2).flatten(1) on (n, 1, npf / 2) tensors, act.py:494-502 -- NOT interleaved)
            const int half = npf >> 1;
            const bool is_cos = jj >= half;
            const float a = (axis == 0 ? x : (axis == 1 ? y : z)) / __ldg(dim_t + (is_cos ? 2 * (jj - half) + 1 : 2 * jj));
            v[u] = is_cos ? cosf(a) : sinf(a);
        }
        *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// rows 0 .. head-1 of the (S, B, E) token tensor: row 0 = latent (B, E), rows 1.. = proprio (head-1, B, E)
__global__ void __launch_bounds__(256) fill_head_rows_kernel(const float* __restrict__ latent, const float* __restrict__ proprio,
                                                             const float* __restrict__ pos, int batch, int head, int E,
                                                             float* __restrict__ tok, __nv_bfloat16* __restrict__ tok_b,
                                                             __nv_bfloat16* __restrict__ tok_pb) {
    const long total = (long)head * batch * E;
    for (long e = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < total; e += (long)gridDim.x * blockDim.x * 4) {
        const long r = e / E;
        const int c = (int)(e - r * E);
        const int s = (int)(r / batch), b = (int)(r % batch);
        const float* src = s == 0 ? latent + (size_t)b * E + c : proprio + ((size_t)(s - 1) * batch + b) * E + c;
        const float4 v = *reinterpret_cast<const float4*>(src);
        *reinterpret_cast<float4*>(tok + e) = v;
        if (tok_b) *reinterpret_cast<uint2*>(tok_b + e) = pack_bf16x4(v.x, v.y, v.z, v.w);
        if (tok_pb) {
            const float4 p = *reinterpret_cast<const float4*>(pos + e);
            *reinterpret_cast<uint2*>(tok_pb + e) = pack_bf16x4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// heads + loss
// ---------------------------------------------------------------------------------------------------------------
constexpr int HL_MAX_OUT = 16;   // action_dim + 1 (is_pad head) <= 16
constexpr int HL_MAX_E = 1024;

struct HeadsLossParams {
    const float* hs; long ld_b, ld_q;          // decoder output row (b, q) at hs + b * ld_b + q * ld_q
    const float* Wa; const float* ba;          // (A, E), (A)
    const float* Wp; const float* bp;          // (1, E), (1)
    const float* actions;                      // (B, Q, A) or NULL (inference: no loss)
    const unsigned char* is_pad;               // (B, Q) bool
    const float* mu; const float* logvar;      // (B, L) or NULL
    int B, Q, E, A, L, sig_start;              // outputs d >= sig_start pass through a sigmoid (RLBench gripper / collision)
    float w_pos; int n_pos;                    // loss weight of the first n_pos dims (position_loss_weight), 1 elsewhere
    float kl_weight;
    float* a_hat; float* is_pad_hat;           // (B, Q, A), (B, Q, 1)
    float* losses;                             // [loss, action_loss, kl_loss]
    double* acc; unsigned int* ticket;         // self-cleaning workspace: acc[1] double, ticket[1]
};

__global__ void __launch_bounds__(256) act_heads_loss_fwd_kernel(HeadsLossParams p) {
    extern __shared__ float sW[];  // (A + 1) x E weights, then A + 1 biases
    const int nout = p.A + 1;
    for (int i = threadIdx.x; i < nout * p.E; i += blockDim.x) sW[i] = i < p.A * p.E ? p.Wa[i] : p.Wp[i - p.A * p.E];
    float* sB = sW + nout * p.E;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) sB[i] = i < p.A ? p.ba[i] : p.bp[0];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long rows = (long)p.B * p.Q;
    float local = 0.f;
    for (long r = (long)blockIdx.x * nwarp + warp; r < rows; r += (long)gridDim.x * nwarp) {
        const int b = (int)(r / p.Q), q = (int)(r % p.Q);
        const float* x = p.hs + b * p.ld_b + q * p.ld_q;
        float dot[HL_MAX_OUT];
#pragma unroll
        for (int d = 0; d < HL_MAX_OUT; ++d) dot[d] = 0.f;
        for (int c = lane * 4; c < p.E; c += 128) {
            const float4 xv = *reinterpret_cast<const float4*>(x + c);
#pragma unroll
            for (int d = 0; d < HL_MAX_OUT; ++d) {
                if (d < nout) {
                    const float4 w = *reinterpret_cast<const float4*>(sW + d * p.E + c);
                    dot[d] = fmaf(xv.w, w.w, fmaf(xv.z, w.z, fmaf(xv.y, w.y, fmaf(xv.x, w.x, dot[d]))));
                }
            }
        }
#pragma unroll
        for (int d = 0; d < HL_MAX_OUT; ++d) {
            if (d < nout) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot[d] += __shfl_xor_sync(PCM_FULL_MASK, dot[d], o);
            }
        }
        if (lane == 0) {
            const bool pad = p.is_pad ? p.is_pad[r] != 0 : false;
#pragma unroll
            for (int d = 0; d < HL_MAX_OUT; ++d) {
                if (d < p.A) {
                    float v = dot[d] + sB[d];
                    if (d >= p.sig_start) v = 1.f / (1.f + expf(-v));
                    p.a_hat[r * p.A + d] = v;
                    if (p.actions && !pad) {
                        const float diff = v - p.actions[r * p.A + d];
                        local = fmaf(d < p.n_pos ? p.w_pos : 1.f, diff * diff, local);
                    }
                } else if (d == p.A) {
                    p.is_pad_hat[r] = dot[d] + sB[d];
                }
            }
        }
    }
    if (!p.actions) return;
    // block reduction of the masked squared error -> one fp64 atomic per CTA; the last CTA finishes the three losses
    __shared__ float red[8];
    __shared__ bool last;
    if (lane == 0) red[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += red[w];
        atomicAdd(p.acc, (double)s);
        __threadfence();
        last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    // KL(q || N(0, I)): klds = -0.5 * (1 + logvar - mu^2 - exp(logvar)), summed over the latent, mean over the batch
    float kl = 0.f;
    if (p.mu) {
        for (int i = threadIdx.x; i < p.B * p.L; i += blockDim.x) {
            const float m = p.mu[i], lv = p.logvar[i];
            kl += -0.5f * (1.f + lv - m * m - expf(lv));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(PCM_FULL_MASK, kl, o);
    __syncthreads();
    if (lane == 0) red[warp] = kl;
    __syncthreads();
    if (threadIdx.x == 0) {
        float k = 0.f;
        for (int w = 0; w < nwarp; ++w) k += red[w];
        k /= (float)p.B;
        __threadfence();
        const double sum = *reinterpret_cast<volatile double*>(p.acc);
        const float action_loss = (float)(sum / ((double)rows * p.A));
        p.losses[0] = action_loss + (p.mu ? k * p.kl_weight : 0.f);
        p.losses[1] = action_loss;
        p.losses[2] = p.mu ? k : 0.f;
        *p.acc = 0.0;      // leave the workspace clean for the next launch
        *p.ticket = 0u;
    }
}

struct HeadsLossBwdParams {
    HeadsLossParams f;
    const float* g_loss; const float* g_action; const float* g_kl;  // upstream scalars (any may be NULL)
    const float* g_a_hat; const float* g_pad;                       // optional upstream gradients of the head outputs
    float* d_hs; long dld_b, dld_q;                                 // gradient of hs, same (b, q) addressing
    float* dWa; float* dba; float* dWp; float* dbp;                 // ACCUMULATED into (atomics); dWp / dbp may be NULL
    float* dmu; float* dlogvar;                                     // (B, L) or NULL
};

constexpr int HL_CHUNK = 128;  // rows whose head gradients are staged in shared memory at a time

__global__ void __launch_bounds__(256) act_heads_loss_bwd_kernel(HeadsLossBwdParams p) {
    extern __shared__ float sW[];  // (A + 1) x E head weights
    __shared__ float sDp[HL_CHUNK][HL_MAX_OUT];  // d(loss)/d(pre-activation) of the chunk's rows
    __shared__ float sdb[HL_MAX_OUT];
    const HeadsLossParams& f = p.f;
    const int nout = f.A + 1;
    for (int i = threadIdx.x; i < nout * f.E; i += blockDim.x) sW[i] = i < f.A * f.E ? f.Wa[i] : f.Wp[i - f.A * f.E];
    if (threadIdx.x < HL_MAX_OUT) sdb[threadIdx.x] = 0.f;
    const float c_act = (p.g_loss ? *p.g_loss : 0.f) + (p.g_action ? *p.g_action : 0.f);
    const float c_kl = (p.g_loss ? *p.g_loss * f.kl_weight : 0.f) + (p.g_kl ? *p.g_kl : 0.f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long rows = (long)f.B * f.Q;
    const float inv = 2.f / (float)((double)rows * f.A);
    // weight-gradient ownership: thread owns the channel quad cq of row group grp (rows rr = grp mod ngrp of a chunk)
    const int quads = f.E / 4, ngrp = blockDim.x / quads > 0 ? blockDim.x / quads : 1;
    const int cq = (threadIdx.x % quads) * 4, grp = threadIdx.x / quads;
    const bool owner = grp < ngrp;
    float dw[HL_MAX_OUT][4];
#pragma unroll
    for (int d = 0; d < HL_MAX_OUT; ++d) dw[d][0] = dw[d][1] = dw[d][2] = dw[d][3] = 0.f;
    const long per_cta = (rows + gridDim.x - 1) / gridDim.x;
    const long r_begin = blockIdx.x * per_cta, r_end = r_begin + per_cta < rows ? r_begin + per_cta : rows;
    for (long r0 = r_begin; r0 < r_end; r0 += HL_CHUNK) {
        const int n = (int)(r_end - r0 < HL_CHUNK ? r_end - r0 : HL_CHUNK);
        __syncthreads();
        // phase 1: head gradients of the chunk (one thread per row)
        if (threadIdx.x < n) {
            const long r = r0 + threadIdx.x;
            const bool pad = f.is_pad ? f.is_pad[r] != 0 : false;
#pragma unroll
            for (int d = 0; d < HL_MAX_OUT; ++d) {
                float g = 0.f;
                if (d < f.A) {
                    const float v = f.a_hat[r * f.A + d];
                    g = p.g_a_hat ? p.g_a_hat[r * f.A + d] : 0.f;
                    if (f.actions && !pad) g = fmaf(c_act * inv * (d < f.n_pos ? f.w_pos : 1.f), v - f.actions[r * f.A + d], g);
                    if (d >= f.sig_start) g *= v * (1.f - v);
                } else if (d == f.A) {
                    g = p.g_pad ? p.g_pad[r] : 0.f;
                }
                sDp[threadIdx.x][d] = g;
                if (d < nout && g != 0.f) atomicAdd(&sdb[d], g);
            }
        }
        __syncthreads();
        // phase 2: d(hs) rows = dpre . W, one warp per row
        for (int rr = warp; rr < n; rr += nwarp) {
            const long r = r0 + rr;
            const int b = (int)(r / f.Q), q = (int)(r % f.Q);
            float* dx = p.d_hs + b * p.dld_b + q * p.dld_q;
            for (int c = lane * 4; c < f.E; c += 128) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int d = 0; d < HL_MAX_OUT; ++d) {
                    if (d < nout) {
                        const float g = sDp[rr][d];
                        const float4 w = *reinterpret_cast<const float4*>(sW + d * f.E + c);
                        acc.x = fmaf(g, w.x, acc.x); acc.y = fmaf(g, w.y, acc.y);
                        acc.z = fmaf(g, w.z, acc.z); acc.w = fmaf(g, w.w, acc.w);
                    }
                }
                *reinterpret_cast<float4*>(dx + c) = acc;
            }
        }
        // phase 3: dW[d, cq..cq+3] += sum_rows dpre[row, d] * hs[row, cq..cq+3] (registers; rows strided over groups)
        if (owner) {
            // independent iterations: unrolled so that several rows' loads are in flight (a serial chain of L2 / HBM latencies
            // was 100+ us of this kernel with one load outstanding per thread)
#pragma unroll 8
            for (int rr = grp; rr < n; rr += ngrp) {
                const long r = r0 + rr;
                const int b = (int)(r / f.Q), q = (int)(r % f.Q);
                const float4 xv = __ldg(reinterpret_cast<const float4*>(f.hs + b * f.ld_b + q * f.ld_q + cq));
#pragma unroll
                for (int d = 0; d < HL_MAX_OUT; ++d) {
                    if (d < nout) {
                        const float g = sDp[rr][d];
                        dw[d][0] = fmaf(g, xv.x, dw[d][0]); dw[d][1] = fmaf(g, xv.y, dw[d][1]);
                        dw[d][2] = fmaf(g, xv.z, dw[d][2]); dw[d][3] = fmaf(g, xv.w, dw[d][3]);
                    }
                }
            }
        }
    }
    // combine the row groups of the CTA in shared memory (the weight tile is no longer needed), then ONE global atomic per
    // weight element per CTA: 24 CTAs x (A + 1) x E atomics instead of one per owner thread
    __syncthreads();
    for (int i = threadIdx.x; i < nout * f.E; i += blockDim.x) sW[i] = 0.f;
    __syncthreads();
    if (owner) {
#pragma unroll
        for (int d = 0; d < HL_MAX_OUT; ++d) {
            if (d < nout) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (dw[d][u] != 0.f) atomicAdd(&sW[d * f.E + cq + u], dw[d][u]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nout * f.E; i += blockDim.x) {
        const float v = sW[i];
        if (v != 0.f) {
            if (i < f.A * f.E) atomicAdd(p.dWa + i, v);
            else if (p.dWp) atomicAdd(p.dWp + (i - f.A * f.E), v);
        }
    }
    if (threadIdx.x < nout && sdb[threadIdx.x] != 0.f) {
        if (threadIdx.x < f.A) atomicAdd(p.dba + threadIdx.x, sdb[threadIdx.x]);
        else if (p.dbp) atomicAdd(p.dbp, sdb[threadIdx.x]);
    }
    if (blockIdx.x == 0 && f.mu && p.dmu) {
        const float s = c_kl / (float)f.B;
        for (int i = threadIdx.x; i < f.B * f.L; i += blockDim.x) {
            p.dmu[i] = s * f.mu[i];
            p.dlogvar[i] = s * 0.5f * (expf(f.logvar[i]) - 1.f);
        }
    }
}

}  // namespace

PCM_API int pcm_coord_embed_sine_tokens(int per_cloud, int batch, int head_rows, int E, int npf, const float* coord,
                                        const float* dim_t, const float* add_pos, float* pos, pcm_stream_t stream) {
    if (per_cloud <= 0 || batch <= 0) return PCM_OK;
    if (!coord || !dim_t || !pos || (head_rows > 0 && !add_pos)) return PCM_EINVAL;
    if (E % 4 || E > 4096 || npf <= 0 || (npf & 1) || 3 * npf > E || head_rows < 0) return PCM_EUNSUPPORTED;
    const long rows = (long)(head_rows + per_cloud) * batch;
    const int threads = E / 4 >= 256 ? 256 : ((E / 4 + 31) / 32) * 32;
    coord_embed_sine_tokens_kernel<<<(unsigned)rows, threads, 0, pcm_cu_stream(stream)>>>(coord, dim_t, add_pos, per_cloud, batch, head_rows,
                                                                                        E, npf, pos);
    return pcm_launch_status();
}

PCM_API int pcm_fill_head_rows(int batch, int head_rows, int E, const float* latent, const float* proprio, const float* pos,
                               float* tokens, void* tokens_bf16, void* tokens_pos_bf16, pcm_stream_t stream) {
    if (batch <= 0 || head_rows <= 0) return PCM_OK;
    if (!latent || (head_rows > 1 && !proprio) || !tokens || (tokens_pos_bf16 && !pos)) return PCM_EINVAL;
    if (E % 4) return PCM_EUNSUPPORTED;
    const long total = (long)head_rows * batch * E;
    fill_head_rows_kernel<<<pcm_divup(total / 4, 256), 256, 0, pcm_cu_stream(stream)>>>(
        latent, proprio, pos, batch, head_rows, E, tokens, reinterpret_cast<__nv_bfloat16*>(tokens_bf16),
        reinterpret_cast<__nv_bfloat16*>(tokens_pos_bf16));
    return pcm_launch_status();
}

static int heads_check(const HeadsLossParams& p) {
    if (!p.hs || !p.Wa || !p.ba || !p.Wp || !p.bp || !p.a_hat || !p.is_pad_hat) return PCM_EINVAL;
    if (p.actions && (!p.losses || !p.acc || !p.ticket)) return PCM_EINVAL;
    if ((p.mu == nullptr) != (p.logvar == nullptr)) return PCM_EINVAL;
    if (p.E % 128 || p.E > HL_MAX_E || p.A <= 0 || p.A + 1 > HL_MAX_OUT || p.B <= 0 || p.Q <= 0) return PCM_EUNSUPPORTED;
    if (p.ld_b % 4 || p.ld_q % 4) return PCM_EUNSUPPORTED;
    return PCM_OK;
}

PCM_API int pcm_act_heads_loss_fwd(int B, int Q, int E, int A, int L, int sig_start, int n_pos, float w_pos, float kl_weight,
                                   const float* hs, long long ld_b, long long ld_q, const float* Wa, const float* ba,
                                   const float* Wp, const float* bp, const float* actions, const unsigned char* is_pad,
                                   const float* mu, const float* logvar, float* a_hat, float* is_pad_hat, float* losses,
                                   double* acc, unsigned int* ticket, pcm_stream_t stream) {
    HeadsLossParams p{hs, (long)ld_b, (long)ld_q, Wa, ba, Wp, bp, actions, is_pad, mu, logvar, B, Q, E, A, L, sig_start, w_pos, n_pos,
                      kl_weight, a_hat, is_pad_hat, losses, acc, ticket};
    const int st = heads_check(p);
    if (st != PCM_OK) return st;
    const size_t smem = ((size_t)(A + 1) * E + (A + 1)) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(act_heads_loss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        attr = true;
    }
    const long rows = (long)B * Q;
    // latency-bound (a warp walks its rows one after the other): four CTAs per SM keep enough loads in flight
    // (6 400 rows: 44 us at 148 CTAs, 30 at 296, 25 at 592, 30 at 800)
    const int grid = (int)(rows / 8 < 592 ? (rows + 7) / 8 : 592);
    act_heads_loss_fwd_kernel<<<grid, 256, smem, pcm_cu_stream(stream)>>>(p);
    return pcm_launch_status();
}

PCM_API int pcm_act_heads_loss_bwd(int B, int Q, int E, int A, int L, int sig_start, int n_pos, float w_pos, float kl_weight,
                                   const float* hs, long long ld_b, long long ld_q, const float* Wa, const float* Wp,
                                   const float* actions, const unsigned char* is_pad, const float* mu, const float* logvar,
                                   const float* a_hat, const float* g_loss, const float* g_action, const float* g_kl,
                                   const float* g_a_hat, const float* g_pad, float* d_hs, long long dld_b, long long dld_q,
                                   float* dWa, float* dba, float* dWp, float* dbp, float* dmu, float* dlogvar,
                                   pcm_stream_t stream) {
    HeadsLossBwdParams p{};
    p.f = HeadsLossParams{hs, (long)ld_b, (long)ld_q, Wa, nullptr, Wp, nullptr, actions, is_pad, mu, logvar, B, Q, E, A, L, sig_start,
                          w_pos, n_pos, kl_weight, const_cast<float*>(a_hat), nullptr, nullptr, nullptr, nullptr};
    if (!hs || !Wa || !Wp || !a_hat || !d_hs || !dWa || !dba) return PCM_EINVAL;
    if (E % 128 || E > HL_MAX_E || A <= 0 || A + 1 > HL_MAX_OUT || B <= 0 || Q <= 0 || ld_b % 4 || ld_q % 4 || dld_b % 4 || dld_q % 4)
        return PCM_EUNSUPPORTED;
    if ((g_pad != nullptr) && (!dWp || !dbp)) return PCM_EINVAL;
    p.g_loss = g_loss; p.g_action = g_action; p.g_kl = g_kl; p.g_a_hat = g_a_hat; p.g_pad = g_pad;
    p.d_hs = d_hs; p.dld_b = (long)dld_b; p.dld_q = (long)dld_q;
    p.dWa = dWa; p.dba = dba; p.dWp = dWp; p.dbp = dbp; p.dmu = dmu; p.dlogvar = dlogvar;
    const size_t smem = (size_t)(A + 1) * E * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(act_heads_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        attr = true;
    }
    const long rows = (long)B * Q;
    // 32 rows per CTA at least (register-accumulated weight gradients, one combine per CTA); 6 400 rows: 70 us at 74 CTAs,
    // 44 at 148, 40 at 200
    const int grid = (int)(rows / 32 < 296 ? (rows + 31) / 32 : 296);
    act_heads_loss_bwd_kernel<<<grid, 256, smem, pcm_cu_stream(stream)>>>(p);
    return pcm_launch_status();
}
