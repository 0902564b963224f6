// optimizer.cu -- fused clip-by-global-norm + AdamW over the flat fp32 parameter buffer, sm_100a.
//
// Replaces, for the reference training step, torch.nn.utils.clip_grad_norm_ (Lightning
// `gradient_clip_val: 0.5`, configs/trainer/ddp.yaml:12) followed by torch.optim.AdamW
// (configs/model/maniskill2_act_pcd_model.yaml:11-14) and the per-step OneCycleLR values: two
// launches over contiguous memory instead of ~10 multi-tensor launches over 250 tensors.
//   pass 1: sum of squares of the (already all-reduced) flat gradient -> one fp64 scalar;
//   pass 2: p, m, v update with the clip coefficient computed ON DEVICE from that scalar and
//           lr / beta1 / bias corrections read from a small device-resident `hyper` vector, so
//           the step has no host synchronisation and can live inside a CUDA graph.
// HBM-bound: 16 B read + 12 B (+2 B bf16 operand copy) written per parameter (+4 B read in pass 1): 128-bit accesses,
// grid = a multiple of the SM count with a grid-stride loop.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n4, double* __restrict__ out) {
    float acc = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(PCM_FULL_MASK, acc, o);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = s[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xff, t, o);
        if (threadIdx.x == 0) atomicAdd(out, (double)t);
    }
}

// hyper = [lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2, clip_norm, grad_scale]
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long n4, const float* __restrict__ hyper,
                                                    const double* __restrict__ sumsq, float* __restrict__ norm_out,
                                                    __nv_bfloat16* __restrict__ p_bf16, int zero_grad) {
    // zero_grad: leave the gradient buffer ZEROED instead of writing the clipped gradient back (same 4 B/param of stores):
    // the next step's separate zero-fill pass over the flat gradient (96 MB ACT, 1 GB Diffusion Policy) disappears
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const float bc1 = hyper[5], bc2 = hyper[6], clip = hyper[7], gscale = hyper[8];
    const float norm = (float)sqrt(*sumsq) * gscale;
    float coef = gscale;
    if (clip > 0.f) coef *= fminf(clip / (norm + 1e-6f), 1.0f);  // torch.nn.utils.clip_grad_norm_
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = norm;
    const float decay = 1.0f - lr * wd;
    const float step = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = G[k] * coef;
            G[k] = gk;
            M[k] = M[k] + (1.0f - b1) * (gk - M[k]);                 // exp_avg.lerp_(grad, 1 - beta1)
            V[k] = V[k] * b2 + (1.0f - b2) * gk * gk;                // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
            const float denom = sqrtf(V[k]) * inv_sqrt_bc2 + eps;    // (sqrt(v) / sqrt(bc2)).add_(eps)
            P[k] = P[k] * decay - step * (M[k] / denom);             // p.mul_(1 - lr wd); p.addcdiv_(m, denom, -lr/bc1)
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        if (p_bf16 != nullptr) {  // bf16 operand copy read by the tensor-core GEMMs of the next step
            __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            reinterpret_cast<uint2*>(p_bf16)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
        reinterpret_cast<float4*>(g)[i] = zero_grad ? make_float4(0.f, 0.f, 0.f, 0.f) : gg;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
}

// Slice lists: n equally sized slices that live at unrelated addresses (the same sub-block of every decoder layer's
// in_proj_weight inside the flat parameter / gradient buffers).  ptrs[] is a device array of n addresses.
//   gather: dst[s * elems + i]  = src_s[i]     (stack the slices into one GEMM operand)
//   add:    dst_s[i]           += src[s * elems + i]   (hand a stacked gradient back to the slices)
template <typename T>
__global__ void __launch_bounds__(256) gather_slices_kernel(const long long* __restrict__ ptrs, long elems16, T* __restrict__ dst) {
    const uint4* src = reinterpret_cast<const uint4*>(ptrs[blockIdx.y]);
    uint4* d = reinterpret_cast<uint4*>(dst) + (size_t)blockIdx.y * elems16;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < elems16; i += (long)gridDim.x * blockDim.x) d[i] = src[i];
}

__global__ void __launch_bounds__(256) add_slices_kernel(const long long* __restrict__ ptrs, long elems4, const float* __restrict__ src) {
    float4* d = reinterpret_cast<float4*>(ptrs[blockIdx.y]);
    const float4* s = reinterpret_cast<const float4*>(src) + (size_t)blockIdx.y * elems4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < elems4; i += (long)gridDim.x * blockDim.x) {
        float4 a = d[i];
        const float4 b = s[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        d[i] = a;
    }
}

}  // namespace

// bytes_per_slice % 16 == 0, all addresses 16-byte aligned.
PCM_API int pcm_gather_slices(int n, long long bytes_per_slice, const long long* ptrs, void* dst, pcm_stream_t stream) {
    if (n <= 0 || bytes_per_slice <= 0) return PCM_OK;
    if (!ptrs || !dst) return PCM_EINVAL;
    if ((bytes_per_slice % 16) || (reinterpret_cast<uintptr_t>(dst) & 15)) return PCM_EUNSUPPORTED;
    const long e16 = bytes_per_slice / 16;
    const int gx = (int)((e16 + 255) / 256 < 64 ? (e16 + 255) / 256 : 64);
    gather_slices_kernel<uint4><<<dim3(gx, n), 256, 0, pcm_cu_stream(stream)>>>(ptrs, e16, reinterpret_cast<uint4*>(dst));
    return pcm_launch_status();
}

PCM_API int pcm_add_slices(int n, long long floats_per_slice, const long long* ptrs, const float* src, pcm_stream_t stream) {
    if (n <= 0 || floats_per_slice <= 0) return PCM_OK;
    if (!ptrs || !src) return PCM_EINVAL;
    if ((floats_per_slice % 4) || (reinterpret_cast<uintptr_t>(src) & 15)) return PCM_EUNSUPPORTED;
    const long e4 = floats_per_slice / 4;
    const int gx = (int)((e4 + 255) / 256 < 64 ? (e4 + 255) / 256 : 64);
    add_slices_kernel<<<dim3(gx, n), 256, 0, pcm_cu_stream(stream)>>>(ptrs, e4, src);
    return pcm_launch_status();
}

// n must be a multiple of 4 and the buffers 16-byte aligned (FlatState pads every tensor to 4).
// sumsq: one zero-initialised fp64 scratch scalar (re-zeroed by this call for the next step).
PCM_API int pcm_clip_adamw_step_bf16(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                     const float* hyper, double* sumsq, float* norm_out, void* param_bf16,
                                     pcm_stream_t stream) {
    return pcm_clip_adamw_step_ex(n, param, grad, exp_avg, exp_avg_sq, hyper, sumsq, norm_out, param_bf16, 0, stream);
}

// zero_grad = 1: the gradient buffer is left zeroed (ready for the next step's accumulation) instead of holding the
// clipped gradient (what torch.nn.utils.clip_grad_norm_ leaves behind).
PCM_API int pcm_clip_adamw_step_ex(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                   const float* hyper, double* sumsq, float* norm_out, void* param_bf16, int zero_grad,
                                   pcm_stream_t stream) {
    if (n <= 0) return PCM_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper || !sumsq) return PCM_EINVAL;
    if (n % 4) return PCM_EUNSUPPORTED;
    cudaStream_t st = pcm_cu_stream(stream);
    const long n4 = n / 4;
    long blocks = (n4 + 255) / 256;
    const int grid = (int)(blocks < 148L * 8 ? blocks : 148L * 8);
    cudaError_t e = cudaMemsetAsync(sumsq, 0, sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    sumsq_kernel<<<grid, 256, 0, st>>>(grad, n4, sumsq);
    int r = pcm_launch_status();
    if (r) return r;
    adamw_kernel<<<grid, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n4, hyper, sumsq, norm_out,
                                       reinterpret_cast<__nv_bfloat16*>(param_bf16), zero_grad);
    return pcm_launch_status();
}

PCM_API int pcm_clip_adamw_step(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                const float* hyper, double* sumsq, float* norm_out, pcm_stream_t stream) {
    return pcm_clip_adamw_step_bf16(n, param, grad, exp_avg, exp_avg_sq, hyper, sumsq, norm_out, nullptr, stream);
}
