// pointops_misc.cu -- grouping / interpolation / aggregation / subtraction / scatter-attention
// operators of the `pointops` API for sm_100a.  None of them is on the end-to-end training step
// of the reference (zero call sites in src/), they exist so that the
// drop-in `pointops` package is complete.  All are HBM-bound gather / scatter kernels: threads
// are mapped with the channel index fastest so that every gathered row is read as one coalesced
// segment, index / weight values are read once per row through the read-only path, and the
// scatter-attention steps reduce over channels inside the thread (or warp) before touching
// global memory, removing the one-atomic-per-product pattern of the reference.
#include "common.cuh"

namespace {

constexpr int T = 256;

// ---- grouping (reference grouping_cuda_kernel.cu:5-25) ------------------------------------
__global__ void __launch_bounds__(T) grouping_fwd_kernel(long total, int nsample, int c,
                                                         const float* __restrict__ input,
                                                         const int* __restrict__ idx,
                                                         float* __restrict__ output) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long row = index / c;  // m_idx * nsample + nsample_idx
        const int c_idx = (int)(index - row * c);
        output[index] = __ldg(input + (long)__ldg(idx + row) * c + c_idx);
    }
}

__global__ void __launch_bounds__(T) grouping_bwd_kernel(long total, int nsample, int c,
                                                         const float* __restrict__ grad_output,
                                                         const int* __restrict__ idx,
                                                         float* __restrict__ grad_input) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long row = index / c;
        const int c_idx = (int)(index - row * c);
        atomicAdd(grad_input + (long)__ldg(idx + row) * c + c_idx, grad_output[index]);
    }
}

// ---- interpolation (reference interpolation_cuda_kernel.cu:5-33) --------------------------
__global__ void __launch_bounds__(T) interpolation_fwd_kernel(long total, int c, int k,
                                                              const float* __restrict__ input,
                                                              const int* __restrict__ idx,
                                                              const float* __restrict__ weight,
                                                              float* __restrict__ output) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long n_idx = index / c;
        const int c_idx = (int)(index - n_idx * c);
        float acc = output[index];
        for (int i = 0; i < k; i++) {
            const long ii = n_idx * k + i;
            acc = __fmaf_rn(__ldg(input + (long)__ldg(idx + ii) * c + c_idx), __ldg(weight + ii), acc);
        }
        output[index] = acc;
    }
}

__global__ void __launch_bounds__(T) interpolation_bwd_kernel(long total, int c, int k,
                                                              const float* __restrict__ grad_output,
                                                              const int* __restrict__ idx,
                                                              const float* __restrict__ weight,
                                                              float* __restrict__ grad_input) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long n_idx = index / c;
        const int c_idx = (int)(index - n_idx * c);
        const float g = grad_output[index];
        for (int i = 0; i < k; i++) {
            const long ii = n_idx * k + i;
            atomicAdd(grad_input + (long)__ldg(idx + ii) * c + c_idx, __fmul_rn(g, __ldg(weight + ii)));
        }
    }
}

// ---- aggregation (reference aggregation_cuda_kernel.cu:5-39) ------------------------------
__global__ void __launch_bounds__(T) aggregation_fwd_kernel(long total, int nsample, int c, int w_c,
                                                            const float* __restrict__ input,
                                                            const float* __restrict__ position,
                                                            const float* __restrict__ weight,
                                                            const int* __restrict__ idx,
                                                            float* __restrict__ output) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long n_idx = index / c;
        const int c_idx = (int)(index - n_idx * c);
        const int w_c_idx = c_idx % w_c;
        float acc = output[index];
        for (int s = 0; s < nsample; s++) {
            const long ii = n_idx * nsample + s;
            const float v = __fadd_rn(__ldg(input + (long)__ldg(idx + ii) * c + c_idx), __ldg(position + ii * c + c_idx));
            acc = __fmaf_rn(v, __ldg(weight + ii * w_c + w_c_idx), acc);
        }
        output[index] = acc;
    }
}

__global__ void __launch_bounds__(T) aggregation_bwd_kernel(long total, int nsample, int c, int w_c,
                                                            const float* __restrict__ input,
                                                            const float* __restrict__ position,
                                                            const float* __restrict__ weight,
                                                            const int* __restrict__ idx,
                                                            const float* __restrict__ grad_output,
                                                            float* __restrict__ grad_input,
                                                            float* __restrict__ grad_position,
                                                            float* __restrict__ grad_weight) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long n_idx = index / c;
        const int c_idx = (int)(index - n_idx * c);
        const int w_c_idx = c_idx % w_c;
        const float g = grad_output[index];
        for (int s = 0; s < nsample; s++) {
            const long ii = n_idx * nsample + s;
            const long input_idx = (long)__ldg(idx + ii) * c + c_idx;
            const long position_idx = ii * c + c_idx;
            const long weight_idx = ii * w_c + w_c_idx;
            const float w = __ldg(weight + weight_idx);
            atomicAdd(grad_input + input_idx, __fmul_rn(g, w));
            grad_position[position_idx] = __fmul_rn(g, w);
            atomicAdd(grad_weight + weight_idx, __fmul_rn(g, __fadd_rn(__ldg(input + input_idx), __ldg(position + position_idx))));
        }
    }
}

// ---- subtraction (reference subtraction_cuda_kernel.cu:5-30) ------------------------------
__global__ void __launch_bounds__(T) subtraction_fwd_kernel(long total, int nsample, int c,
                                                            const float* __restrict__ input1,
                                                            const float* __restrict__ input2,
                                                            const int* __restrict__ idx,
                                                            float* __restrict__ output) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long row = index / c;  // n_idx * nsample + s
        const int c_idx = (int)(index - row * c);
        const long n_idx = row / nsample;
        output[index] = __fsub_rn(__ldg(input1 + n_idx * c + c_idx), __ldg(input2 + (long)__ldg(idx + row) * c + c_idx));
    }
}

__global__ void __launch_bounds__(T) subtraction_bwd_kernel(long total, int nsample, int c,
                                                            const int* __restrict__ idx,
                                                            const float* __restrict__ grad_output,
                                                            float* __restrict__ grad_input1,
                                                            float* __restrict__ grad_input2) {
    // one thread per (n, c): the nsample contributions to grad_input1 are summed in registers
    // (the reference issues nsample atomics per element); grad_input2 is a true scatter.
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long n_idx = index / c;
        const int c_idx = (int)(index - n_idx * c);
        float acc = 0.f;
        for (int s = 0; s < nsample; ++s) {
            const long row = n_idx * nsample + s;
            const float g = grad_output[row * c + c_idx];
            acc += g;
            atomicAdd(grad_input2 + (long)__ldg(idx + row) * c + c_idx, -g);
        }
        atomicAdd(grad_input1 + index, acc);
    }
}

// ---- scatter attention (reference attention_cuda_kernel.cu:9-86) --------------------------
// relation forward: output[r, g] += sum_c q[t_r, g, c] * k[f_r, g, c] * w[c].  One warp per
// (r, g): lanes stride over c (coalesced), shuffle-reduce, one add.
__global__ void __launch_bounds__(T) attn_relation_fwd_kernel(long pairs, int g, int c,
                                                              const float* __restrict__ query,
                                                              const float* __restrict__ key,
                                                              const float* __restrict__ weight,
                                                              const int* __restrict__ index_target,
                                                              const int* __restrict__ index_refer,
                                                              float* __restrict__ output) {
    const int lane = threadIdx.x & 31;
    for (long w = ((long)blockIdx.x * T + threadIdx.x) >> 5; w < pairs; w += ((long)gridDim.x * T) >> 5) {
        const long r = w / g;
        const int gi = (int)(w - r * g);
        const float* q = query + ((long)__ldg(index_target + r) * g + gi) * c;
        const float* k = key + ((long)__ldg(index_refer + r) * g + gi) * c;
        float acc = 0.f;
        for (int ci = lane; ci < c; ci += 32) acc += __ldg(q + ci) * __ldg(k + ci) * __ldg(weight + ci);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(PCM_FULL_MASK, acc, o);
        if (lane == 0) output[w] += acc;
    }
}

__global__ void __launch_bounds__(T) attn_relation_bwd_kernel(long total, int g, int c,
                                                              const float* __restrict__ query,
                                                              float* __restrict__ grad_query,
                                                              const float* __restrict__ key,
                                                              float* __restrict__ grad_key,
                                                              const float* __restrict__ weight,
                                                              float* __restrict__ grad_weight,
                                                              const int* __restrict__ index_target,
                                                              const int* __restrict__ index_refer,
                                                              const float* __restrict__ grad_output) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long rg = index / c;  // r * g + g_idx
        const int ci = (int)(index - rg * c);
        const long r = rg / g;
        const int gi = (int)(rg - r * g);
        const long q_idx = ((long)__ldg(index_target + r) * g + gi) * c + ci;
        const long k_idx = ((long)__ldg(index_refer + r) * g + gi) * c + ci;
        const float grad_r = __ldg(grad_output + rg);
        const float qv = __ldg(query + q_idx), kv = __ldg(key + k_idx), wv = __ldg(weight + ci);
        atomicAdd(grad_query + q_idx, grad_r * kv * wv);
        atomicAdd(grad_key + k_idx, grad_r * qv * wv);
        atomicAdd(grad_weight + ci, grad_r * kv * qv);
    }
}

__global__ void __launch_bounds__(T) attn_fusion_fwd_kernel(long total, int g, int c,
                                                            const float* __restrict__ weight,
                                                            const float* __restrict__ value,
                                                            const int* __restrict__ index_target,
                                                            const int* __restrict__ index_refer,
                                                            float* __restrict__ output) {
    for (long index = (long)blockIdx.x * T + threadIdx.x; index < total; index += (long)gridDim.x * T) {
        const long rg = index / c;
        const int ci = (int)(index - rg * c);
        const long r = rg / g;
        const int gi = (int)(rg - r * g);
        const long o_idx = ((long)__ldg(index_target + r) * g + gi) * c + ci;
        const long v_idx = ((long)__ldg(index_refer + r) * g + gi) * c + ci;
        atomicAdd(output + o_idx, __ldg(weight + rg) * __ldg(value + v_idx));
    }
}

// fusion backward: grad_weight[r, g] += sum_c grad_out[t_r, g, c] * value[f_r, g, c] (warp
// reduction over c), grad_value[f_r, g, c] += grad_out[t_r, g, c] * weight[r, g] (scatter).
__global__ void __launch_bounds__(T) attn_fusion_bwd_kernel(long pairs, int g, int c,
                                                            const float* __restrict__ weight,
                                                            float* __restrict__ grad_weight,
                                                            const float* __restrict__ value,
                                                            float* __restrict__ grad_value,
                                                            const int* __restrict__ index_target,
                                                            const int* __restrict__ index_refer,
                                                            const float* __restrict__ grad_output) {
    const int lane = threadIdx.x & 31;
    for (long w = ((long)blockIdx.x * T + threadIdx.x) >> 5; w < pairs; w += ((long)gridDim.x * T) >> 5) {
        const long r = w / g;
        const int gi = (int)(w - r * g);
        const long o_base = ((long)__ldg(index_target + r) * g + gi) * c;
        const long v_base = ((long)__ldg(index_refer + r) * g + gi) * c;
        const float wv = __ldg(weight + w);
        float acc = 0.f;
        for (int ci = lane; ci < c; ci += 32) {
            const float go = __ldg(grad_output + o_base + ci);
            acc += go * __ldg(value + v_base + ci);
            atomicAdd(grad_value + v_base + ci, go * wv);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(PCM_FULL_MASK, acc, o);
        if (lane == 0) grad_weight[w] += acc;
    }
}

inline int grid_for(long total) {
    long blocks = (total + T - 1) / T;
    const long cap = 148L * 32;  // persistent-ish grid: a few waves of 148 SMs, grid-stride inside
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

#define PCM_RETURN_IF_EMPTY(total) if ((total) <= 0) return PCM_OK

PCM_API int pcm_grouping_forward(int m, int nsample, int c, const float* input, const int* idx,
                                 float* output, pcm_stream_t stream) {
    const long total = (long)m * nsample * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!input || !idx || !output) return PCM_EINVAL;
    grouping_fwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, nsample, c, input, idx, output);
    return pcm_launch_status();
}

PCM_API int pcm_grouping_backward(int m, int nsample, int c, const float* grad_output, const int* idx,
                                  float* grad_input, pcm_stream_t stream) {
    const long total = (long)m * nsample * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!grad_output || !idx || !grad_input) return PCM_EINVAL;
    grouping_bwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, nsample, c, grad_output, idx, grad_input);
    return pcm_launch_status();
}

PCM_API int pcm_interpolation_forward(int n, int c, int k, const float* input, const int* idx,
                                      const float* weight, float* output, pcm_stream_t stream) {
    const long total = (long)n * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!input || !idx || !weight || !output) return PCM_EINVAL;
    interpolation_fwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, c, k, input, idx, weight, output);
    return pcm_launch_status();
}

PCM_API int pcm_interpolation_backward(int n, int c, int k, const float* grad_output, const int* idx,
                                       const float* weight, float* grad_input, pcm_stream_t stream) {
    const long total = (long)n * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!grad_output || !idx || !weight || !grad_input) return PCM_EINVAL;
    interpolation_bwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, c, k, grad_output, idx, weight, grad_input);
    return pcm_launch_status();
}

PCM_API int pcm_aggregation_forward(int n, int nsample, int c, int w_c, const float* input,
                                    const float* position, const float* weight, const int* idx,
                                    float* output, pcm_stream_t stream) {
    const long total = (long)n * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!input || !position || !weight || !idx || !output || w_c <= 0) return PCM_EINVAL;
    aggregation_fwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, nsample, c, w_c, input, position, weight, idx, output);
    return pcm_launch_status();
}

PCM_API int pcm_aggregation_backward(int n, int nsample, int c, int w_c, const float* input,
                                     const float* position, const float* weight, const int* idx,
                                     const float* grad_output, float* grad_input,
                                     float* grad_position, float* grad_weight, pcm_stream_t stream) {
    const long total = (long)n * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!input || !position || !weight || !idx || !grad_output || !grad_input || !grad_position || !grad_weight || w_c <= 0) return PCM_EINVAL;
    aggregation_bwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, nsample, c, w_c, input, position, weight, idx,
                                                                           grad_output, grad_input, grad_position, grad_weight);
    return pcm_launch_status();
}

PCM_API int pcm_subtraction_forward(int n, int nsample, int c, const float* input1, const float* input2,
                                    const int* idx, float* output, pcm_stream_t stream) {
    const long total = (long)n * nsample * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!input1 || !input2 || !idx || !output) return PCM_EINVAL;
    subtraction_fwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, nsample, c, input1, input2, idx, output);
    return pcm_launch_status();
}

PCM_API int pcm_subtraction_backward(int n, int nsample, int c, const int* idx, const float* grad_output,
                                     float* grad_input1, float* grad_input2, pcm_stream_t stream) {
    const long total = (long)n * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!idx || !grad_output || !grad_input1 || !grad_input2) return PCM_EINVAL;
    subtraction_bwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, nsample, c, idx, grad_output, grad_input1, grad_input2);
    return pcm_launch_status();
}

PCM_API int pcm_attention_relation_step_forward(int m, int g, int c, const float* query, const float* key,
                                                const float* weight, const int* index_target,
                                                const int* index_refer, float* output,
                                                pcm_stream_t stream) {
    const long pairs = (long)m * g;
    if (pairs <= 0 || c <= 0) return PCM_OK;
    if (!query || !key || !weight || !index_target || !index_refer || !output) return PCM_EINVAL;
    attn_relation_fwd_kernel<<<grid_for(pairs * 32), T, 0, pcm_cu_stream(stream)>>>(pairs, g, c, query, key, weight,
                                                                                  index_target, index_refer, output);
    return pcm_launch_status();
}

PCM_API int pcm_attention_relation_step_backward(int m, int g, int c, const float* query,
                                                 float* grad_query, const float* key, float* grad_key,
                                                 const float* weight, float* grad_weight,
                                                 const int* index_target, const int* index_refer,
                                                 const float* grad_output, pcm_stream_t stream) {
    const long total = (long)m * g * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!query || !grad_query || !key || !grad_key || !weight || !grad_weight || !index_target || !index_refer || !grad_output) return PCM_EINVAL;
    attn_relation_bwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, g, c, query, grad_query, key, grad_key,
                                                                             weight, grad_weight, index_target,
                                                                             index_refer, grad_output);
    return pcm_launch_status();
}

PCM_API int pcm_attention_fusion_step_forward(int m, int g, int c, const float* weight, const float* value,
                                              const int* index_target, const int* index_refer,
                                              float* output, pcm_stream_t stream) {
    const long total = (long)m * g * c;
    PCM_RETURN_IF_EMPTY(total);
    if (!weight || !value || !index_target || !index_refer || !output) return PCM_EINVAL;
    attn_fusion_fwd_kernel<<<grid_for(total), T, 0, pcm_cu_stream(stream)>>>(total, g, c, weight, value, index_target,
                                                                           index_refer, output);
    return pcm_launch_status();
}

PCM_API int pcm_attention_fusion_step_backward(int m, int g, int c, const float* weight,
                                               float* grad_weight, const float* value, float* grad_value,
                                               const int* index_target, const int* index_refer,
                                               const float* grad_output, pcm_stream_t stream) {
    const long pairs = (long)m * g;
    if (pairs <= 0 || c <= 0) return PCM_OK;
    if (!weight || !grad_weight || !value || !grad_value || !index_target || !index_refer || !grad_output) return PCM_EINVAL;
    attn_fusion_bwd_kernel<<<grid_for(pairs * 32), T, 0, pcm_cu_stream(stream)>>>(pairs, g, c, weight, grad_weight, value,
                                                                                grad_value, index_target, index_refer,
                                                                                grad_output);
    return pcm_launch_status();
}

PCM_API int pcm_abi_version(void) { return 1; }
PCM_API const char* pcm_build_info(void) { return "pcm_b200 sm_100a nvcc " __DATE__ " " __TIME__; }
