// frame_filter.cu -- raw camera frames -> packed per-sample point clouds on the GPU (the stage in front of
// the voxel-grid subsampling of grid_sample.cu).
//
// Replaces, for a whole batch at once, the numpy boolean-mask selections the reference's dataset classes run per sample
// in the loader workers:
//   mode 0, ManiSkill2 (src/data/components/maniskill2/maniskill2_single_task_pcd_act.py:196-224): points (P, 4) = xyzw of
//           the selected cameras (128 x 128 pixels each), optional random crop that zeroes everything outside a 112 x 112
//           pixel window (:200-208), keep w > 0, then z > 0.005 (ground removed) or x > -0.8 (include_ground);
//   mode 1, RLBench (src/data/components/rlbench/rlbench_single_task_act.py:266-290): the camera point maps stacked and
//           flattened camera-major, keep the points strictly inside SCENE_BOUNDS (comparison in float64: the reference
//           casts the maps with .astype(float)); optional per-point instance mask: listed invalid ids -> 0, > 0 -> 1 (:291-295).
// numpy's `a[mask]` keeps the survivors in ascending index order: so does this (flag -> chunk counts -> exclusive scan ->
// ordered scatter).  Integer / byte work, HBM-bound: 2 kernels, each point read once per kernel.
#include "common.cuh"

namespace {

constexpr int FF_CHUNK = 1024;  // points per CTA (4 per thread)

struct FrameParams {
    int mode;          // 0 ManiSkill2, 1 RLBench
    int xyz_stride;    // floats per input point (4: xyzw, 3: xyz)
    int include_ground;
    double bounds[6];  // mode 1: xmin, ymin, zmin, xmax, ymax, zmax
    int cam_w, cam_h;  // mode 0 crop: pixels per camera row / rows per camera (index = cam * h * w + row * w + col)
    int crop_size;
};

__device__ __forceinline__ bool frame_keep(const FrameParams& fp, const float* __restrict__ pt, long local, const int* __restrict__ crop) {
    const float x = pt[0], y = pt[1], z = pt[2];
    if (fp.mode == 0) {
        if (!(pt[3] > 0.f)) return false;
        if (crop) {  // the reference zeroes the points outside the window, which then fail w > 0; coords[:, :x0] indexes ROWS
            const long pix = local % ((long)fp.cam_w * fp.cam_h);
            const int r = (int)(pix / fp.cam_w), c = (int)(pix % fp.cam_w);
            if (r < crop[0] || r >= crop[0] + fp.crop_size || c < crop[1] || c >= crop[1] + fp.crop_size) return false;
        }
        return fp.include_ground ? (x > -0.8f) : (z > 0.005f);
    }
    const double dx = (double)x, dy = (double)y, dz = (double)z;
    return dx > fp.bounds[0] && dx < fp.bounds[3] && dy > fp.bounds[1] && dy < fp.bounds[4] && dz > fp.bounds[2] && dz < fp.bounds[5];
}

// pass 1: survivors per chunk (sample-major chunk index)
__global__ void __launch_bounds__(256) frame_count_kernel(const FrameParams fp, const float* __restrict__ xyz, long P, int chunks,
                                                          const int* __restrict__ crop, int* __restrict__ chunk_count) {
    __shared__ int s_cnt[8];
    const int sample = blockIdx.y, chunk = blockIdx.x;
    const float* base = xyz + (size_t)sample * P * fp.xyz_stride;
    const int* cr = crop ? crop + 2 * sample : nullptr;
    int mine = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long i = (long)chunk * FF_CHUNK + u * 256 + threadIdx.x;
        if (i < P && frame_keep(fp, base + (size_t)i * fp.xyz_stride, i, cr)) ++mine;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(PCM_FULL_MASK, mine, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        chunk_count[(size_t)sample * chunks + chunk] = t;
    }
}

// pass 2: ordered scatter.  chunk_base = exclusive scan of chunk_count over the whole batch (sample-major).  Within a chunk
// the order is ascending point index: the chunk is walked in four 256-point rows, each compacted with ballot prefixes.
template <bool COLOR_U8>
__global__ void __launch_bounds__(256) frame_scatter_kernel(const FrameParams fp, const float* __restrict__ xyz, const void* __restrict__ color,
                                                            int color_ch, const float* __restrict__ seg, const float* __restrict__ invalid,
                                                            int n_invalid, long P, int chunks, const int* __restrict__ crop,
                                                            const long long* __restrict__ chunk_base, float* __restrict__ out_xyz,
                                                            float* __restrict__ out_color, int out_ch) {
    __shared__ int s_warp[8];
    __shared__ int s_row_total;
    const int sample = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* base = xyz + (size_t)sample * P * fp.xyz_stride;
    const int* cr = crop ? crop + 2 * sample : nullptr;
    long long pos = chunk_base[(size_t)sample * chunks + chunk];
    for (int u = 0; u < 4; ++u) {
        const long i = (long)chunk * FF_CHUNK + u * 256 + threadIdx.x;
        const bool keep = i < P && frame_keep(fp, base + (size_t)i * fp.xyz_stride, i, cr);
        const unsigned bal = __ballot_sync(PCM_FULL_MASK, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; s_warp[w] = run; run += c; }
            s_row_total = run;
        }
        __syncthreads();
        if (keep) {
            const long long o = pos + s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
            const float* pt = base + (size_t)i * fp.xyz_stride;
            out_xyz[o * 3 + 0] = pt[0]; out_xyz[o * 3 + 1] = pt[1]; out_xyz[o * 3 + 2] = pt[2];
            const size_t src = ((size_t)sample * P + i) * color_ch;
            for (int c = 0; c < color_ch; ++c)
                out_color[o * out_ch + c] = COLOR_U8 ? (float)reinterpret_cast<const unsigned char*>(color)[src + c]
                                                     : reinterpret_cast<const float*>(color)[src + c];
            if (seg) {  // instance-id map -> {0, 1}: listed invalid ids and non-positive ids are background
                float m = seg[(size_t)sample * P + i];
                for (int k = 0; k < n_invalid; ++k)
                    if (m == invalid[k]) m = 0.f;
                out_color[o * out_ch + color_ch] = m > 0.f ? 1.f : m;
            }
        }
        pos += s_row_total;
        __syncthreads();
    }
}

}  // namespace

// Pass 1 of the frame filter: chunk_count[(sample, chunk)] = number of surviving points among the chunk's 1024 (chunks =
// ceil(P / 1024) per sample).  mode 0 (ManiSkill2): xyz_stride 4 (xyzw), crop = NULL or (b, 2) int32 (first row, first column of
// the kept crop_size x crop_size window of every cam_h x cam_w camera image); mode 1 (RLBench): xyz_stride 3, bounds =
// (xmin, ymin, zmin, xmax, ymax, zmax) compared in float64.
PCM_API int pcm_frame_filter_count(int b, long long P, int mode, const float* xyz, int xyz_stride, int include_ground,
                                   const double* bounds, const int* crop, int cam_h, int cam_w, int crop_size, int* chunk_count,
                                   pcm_stream_t stream) {
    if (b <= 0 || P <= 0) return PCM_OK;
    if (!xyz || !chunk_count || (mode != 0 && mode != 1) || (mode == 0 && xyz_stride != 4) || (mode == 1 && (xyz_stride < 3 || !bounds)))
        return PCM_EINVAL;
    if (crop && (cam_h <= 0 || cam_w <= 0 || crop_size <= 0 || P % ((long long)cam_h * cam_w))) return PCM_EINVAL;
    FrameParams fp{};
    fp.mode = mode; fp.xyz_stride = xyz_stride; fp.include_ground = include_ground;
    for (int i = 0; i < 6; ++i) fp.bounds[i] = bounds ? bounds[i] : 0.0;
    fp.cam_w = cam_w; fp.cam_h = cam_h; fp.crop_size = crop_size;
    const int chunks = (int)((P + FF_CHUNK - 1) / FF_CHUNK);
    frame_count_kernel<<<dim3(chunks, b), 256, 0, pcm_cu_stream(stream)>>>(fp, xyz, (long)P, chunks, crop, chunk_count);
    return pcm_launch_status();
}

// Pass 2: chunk_base = exclusive prefix sum of chunk_count over the batch (sample-major, int64).  Writes the survivors in
// ascending (sample, point index) order: out_xyz (N, 3) fp32 and out_color (N, out_ch) fp32 = the color_ch input channels
// (uint8 or fp32) followed, when `seg` (b, P) fp32 is given, by the binarised instance mask (ids listed in `invalid`
// (n_invalid, fp32, DEVICE pointer) -> 0, other ids > 0 -> 1).  out_ch = color_ch + (seg ? 1 : 0).
PCM_API int pcm_frame_filter_scatter(int b, long long P, int mode, const float* xyz, int xyz_stride, int include_ground,
                                     const double* bounds, const int* crop, int cam_h, int cam_w, int crop_size, const void* color,
                                     int color_is_u8, int color_ch, const float* seg, const float* invalid, int n_invalid,
                                     const long long* chunk_base, float* out_xyz, float* out_color, pcm_stream_t stream) {
    if (b <= 0 || P <= 0) return PCM_OK;
    if (!xyz || !chunk_base || !out_xyz || (color_ch > 0 && (!color || !out_color)) || (seg && !out_color) || (n_invalid > 0 && !invalid))
        return PCM_EINVAL;
    if ((mode != 0 && mode != 1) || (mode == 0 && xyz_stride != 4) || (mode == 1 && (xyz_stride < 3 || !bounds))) return PCM_EINVAL;
    if (crop && (cam_h <= 0 || cam_w <= 0 || crop_size <= 0 || P % ((long long)cam_h * cam_w))) return PCM_EINVAL;
    FrameParams fp{};
    fp.mode = mode; fp.xyz_stride = xyz_stride; fp.include_ground = include_ground;
    for (int i = 0; i < 6; ++i) fp.bounds[i] = bounds ? bounds[i] : 0.0;
    fp.cam_w = cam_w; fp.cam_h = cam_h; fp.crop_size = crop_size;
    const int chunks = (int)((P + FF_CHUNK - 1) / FF_CHUNK);
    const int out_ch = color_ch + (seg ? 1 : 0);
    const dim3 grid(chunks, b);
    if (color_is_u8)
        frame_scatter_kernel<true><<<grid, 256, 0, pcm_cu_stream(stream)>>>(fp, xyz, color, color_ch, seg, invalid, n_invalid, (long)P, chunks,
                                                                           crop, chunk_base, out_xyz, out_color, out_ch);
    else
        frame_scatter_kernel<false><<<grid, 256, 0, pcm_cu_stream(stream)>>>(fp, xyz, color, color_ch, seg, invalid, n_invalid, (long)P, chunks,
                                                                            crop, chunk_base, out_xyz, out_color, out_ch);
    return pcm_launch_status();
}
