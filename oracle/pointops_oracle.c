/*
 * oracle/pointops_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's `libs/pointops` CUDA kernels
 * (HaoyiZhu/PointCloudMatters @ 3b9df50).  The reference ships NO CPU implementation of these
 * kernels; this file restates each `__global__` function as the sequential program one GPU
 * thread (or one thread block, for FPS) executes, so that index outputs are bit-identical --
 * including tie order, which for kNN / ball query is an artefact of the sequential binary heap.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load this library.  The product path (pointcloudmatters_b200/) never does.
 *
 * Parity pin: the reference has no tests or golden vectors for this path, so
 * this restatement is pinned against the UNMODIFIED reference kernels themselves, compiled from
 * /root/reference into oracle/_ref/libpointops_ref.so (oracle/build_ref.sh) and executed on the
 * B200 by tests/test_parity_ref_gpu.py, and against the committed fixtures in tests/golden/.
 *
 * Distance arithmetic.  nvcc contracts `dx*dx + dy*dy + dz*dz` of the reference into
 * FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.) -- the plain multiply lands on the Y term (read off
 * the sm_100a SASS of the reference build: the FMUL operand is the register loaded from +0x4,
 * and confirmed 100% bit-exact against kNN distances produced by the reference kernels on a
 * B200, tests/golden/):
 *     d = fmaf(dz, dz, fmaf(dx, dx, dy * dy))
 * with dx = x_k - x_last for FPS and dx = q_x - x_k for the query kernels.  Compile with
 * -ffp-contract=off so that gcc never fuses anything we did not write as fmaf().
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* libs/pointops/src/cuda_utils.h:11-14 (opt_n_threads): largest power of two <= work_size,
 * capped at TOTAL_THREADS = 1024 (cuda_utils.h:7), floor 1. */
ORACLE_API int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

static inline float dist2_fps(const float *xyz, int k, float x1, float y1, float z1) {
    /* sampling_cuda_kernel.cu:51-54 */
    const float dx = xyz[k * 3 + 0] - x1;
    const float dy = xyz[k * 3 + 1] - y1;
    const float dz = xyz[k * 3 + 2] - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

static inline float dist2_query(const float *xyz, int i, float qx, float qy, float qz) {
    /* knn_query_cuda_kernel.cu:88-92, ball_query_cuda_kernel.cu:90-94 */
    const float dx = qx - xyz[i * 3 + 0];
    const float dy = qy - xyz[i * 3 + 1];
    const float dz = qz - xyz[i * 3 + 2];
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------------------------
 * Farthest point sampling.
 * Follows farthest_point_sampling_cuda_kernel<block_size> (sampling_cuda_kernel.cu:14-129) and
 * its launcher (:131-171): one block per cloud, block_size = opt_n_threads(n) where n is the
 * caller-supplied maximum cloud size (functions/sampling.py:14-16).  Thread `tid` scans points
 * start_n+tid, +block_size, ... keeping the FIRST strict maximum (:57-58), then the shared-memory
 * tree reduction keeps the lower SLOT on ties (__update, :5-10: `v2 > v1 ? i2 : i1`).  Because the
 * tree folds the upper half onto the lower half first, the surviving thread among equal maxima is
 * the one with the smallest BIT-REVERSED thread id (even beats odd at the last level, ...), not
 * the smallest thread id; the tree is emulated literally below.
 * `tmp` is the caller-initialised running min-distance buffer (1e10, sampling.py:18).
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_farthest_point_sampling(int b, int n, const float *xyz, const int *offset,
                                               const int *new_offset, float *tmp, int *idx) {
    int block_size = oracle_opt_n_threads(n);
    /* launcher switch (:133-170) only instantiates powers of two 1..1024; opt_n_threads only
     * produces those, the `default:` arm (512) is unreachable. */
#pragma omp parallel for schedule(dynamic, 1)
    for (int bid = 0; bid < b; ++bid) {
        float *dists = (float *)malloc(sizeof(float) * block_size);
        int *dists_i = (int *)malloc(sizeof(int) * block_size);
        int start_n = bid == 0 ? 0 : offset[bid - 1];
        int end_n = offset[bid];
        int start_m = bid == 0 ? 0 : new_offset[bid - 1];
        int end_m = new_offset[bid];
        int old = start_n; /* :22-34: old = 0 for bid 0, offset[bid-1] otherwise */
        if (end_m > start_m) idx[start_m] = start_n; /* :39 (guarded: reference writes even when the cloud asks for 0 samples) */
        for (int j = start_m + 1; j < end_m; ++j) {
            const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
            for (int tid = 0; tid < block_size; ++tid) {
                int besti = start_n;
                float best = -1.0f;
                for (int k = start_n + tid; k < end_n; k += block_size) {
                    float d = dist2_fps(xyz, k, x1, y1, z1);
                    float d2 = fminf(d, tmp[k]);
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = block_size / 2; s >= 1; s >>= 1) { /* :64-123 */
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2; /* max(v1, v2) */
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            idx[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* ------------------------------------------------------------------------------------------
 * Binary max-heap helpers -- knn_query_cuda_kernel.cu:15-42 (identical copies in
 * ball_query_cuda_kernel.cu:15-42).
 * ------------------------------------------------------------------------------------------ */
static inline void swap_f(float *x, float *y) { float t = *x; *x = *y; *y = t; }
static inline void swap_i(int *x, int *y) { int t = *x; *x = *y; *y = t; }

static void reheap(float *dist, int *idx, int k) {
    int root = 0;
    int child = root * 2 + 1;
    while (child < k) {
        if (child + 1 < k && dist[child + 1] > dist[child]) child++;
        if (dist[root] > dist[child]) return;
        swap_f(&dist[root], &dist[child]);
        swap_i(&idx[root], &idx[child]);
        root = child;
        child = root * 2 + 1;
    }
}

static void heap_sort(float *dist, int *idx, int k) {
    for (int i = k - 1; i > 0; i--) {
        swap_f(&dist[0], &dist[i]);
        swap_i(&idx[0], &idx[i]);
        reheap(dist, idx, i);
    }
}

/* get_bt_idx, knn_query_cuda_kernel.cu:45-56: first cloud whose new_offset exceeds pt_idx. */
static inline int get_bt_idx(int idx, const int *offset) {
    int i = 0;
    while (1) {
        if (idx < offset[i]) break;
        i++;
    }
    return i;
}

/* ------------------------------------------------------------------------------------------
 * kNN query -- knn_query_cuda_kernel (knn_query_cuda_kernel.cu:60-104): per query a brute-force
 * scan of its cloud, strict `d2 < best_dist[0]` replacement of the heap root, then heap_sort
 * (ascending).  Pads with idx -1 / dist2 1e10 when the cloud has fewer than nsample points.
 * Writes SQUARED distances (functions/query.py:23 applies sqrt afterwards).  nsample <= 128.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_knn_query(int m, int nsample, const float *xyz, const float *new_xyz,
                                 const int *offset, const int *new_offset, int *idx, float *dist2) {
#pragma omp parallel for schedule(static, 64)
    for (int pt = 0; pt < m; ++pt) {
        float best_dist[128];
        int best_idx[128];
        const int bt = get_bt_idx(pt, new_offset);
        const int start = bt == 0 ? 0 : offset[bt - 1];
        const int end = offset[bt];
        const float qx = new_xyz[pt * 3 + 0], qy = new_xyz[pt * 3 + 1], qz = new_xyz[pt * 3 + 2];
        for (int i = 0; i < nsample; i++) { best_dist[i] = 1e10f; best_idx[i] = -1; }
        for (int i = start; i < end; i++) {
            float d2 = dist2_query(xyz, i, qx, qy, qz);
            if (d2 < best_dist[0]) {
                best_dist[0] = d2;
                best_idx[0] = i;
                reheap(best_dist, best_idx, nsample);
            }
        }
        heap_sort(best_dist, best_idx, nsample);
        for (int i = 0; i < nsample; i++) {
            idx[pt * nsample + i] = best_idx[i];
            dist2[pt * nsample + i] = best_dist[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Ball query -- ball_query_cuda_kernel (ball_query_cuda_kernel.cu:58-123).  Quirks restated
 * verbatim: (1) `d2 <= 1e-5` compares in DOUBLE (the literal is a double; SASS has F2F.F64.F32 +
 * DSETP); (2) heap_sort is applied to the scan-ordered candidate array WITHOUT building a heap
 * first (:103), so the result is a deterministic permutation but not in general sorted;
 * (3) when more than nsample candidates exist the strided subsample writes the candidate INDEX,
 * converted to float, into dist2 (:120).  The reference's candidate arrays hold 2048 entries and
 * overflow silently beyond that (undefined behaviour); this restatement stops collecting at 2048
 * -- documented divergence, flagged through *overflow when non-NULL.
 * ------------------------------------------------------------------------------------------ */
#define BALL_MAX_CANDI 2048
ORACLE_API void oracle_ball_query(int m, int nsample, float min_radius, float max_radius,
                                  const float *xyz, const float *new_xyz, const int *offset,
                                  const int *new_offset, int *idx, float *dist2, int *overflow) {
    int any_overflow = 0;
#pragma omp parallel for schedule(static, 16) reduction(| : any_overflow)
    for (int pt = 0; pt < m; ++pt) {
        float *candi_dist = (float *)malloc(sizeof(float) * BALL_MAX_CANDI);
        int *candi_idx = (int *)malloc(sizeof(int) * BALL_MAX_CANDI);
        int candi_num = 0;
        const int bt = get_bt_idx(pt, new_offset);
        const int start = bt == 0 ? 0 : offset[bt - 1];
        const int end = offset[bt];
        const float max_radius2 = max_radius * max_radius;
        const float min_radius2 = min_radius * min_radius;
        const float qx = new_xyz[pt * 3 + 0], qy = new_xyz[pt * 3 + 1], qz = new_xyz[pt * 3 + 2];
        for (int i = start; i < end; i++) {
            float d2 = dist2_query(xyz, i, qx, qy, qz);
            if ((double)d2 <= 1e-5 || (d2 >= min_radius2 && d2 < max_radius2)) {
                if (candi_num >= BALL_MAX_CANDI) { any_overflow = 1; break; }
                candi_dist[candi_num] = d2;
                candi_idx[candi_num] = i;
                candi_num += 1;
            }
        }
        heap_sort(candi_dist, candi_idx, candi_num);
        int *o_idx = idx + (size_t)pt * nsample;
        float *o_d = dist2 + (size_t)pt * nsample;
        if (candi_num <= nsample) {
            for (int i = 0; i < candi_num; i++) { o_idx[i] = candi_idx[i]; o_d[i] = candi_dist[i]; }
            for (int i = candi_num; i < nsample; i++) { o_idx[i] = -1; o_d[i] = 1e10f; }
        } else {
            float sep = (float)candi_num / nsample;
            for (int i = 0; i < nsample; i++) {
                int index = (int)(sep * i);
                o_idx[i] = candi_idx[index];
                o_d[i] = (float)candi_idx[index]; /* :120, sic */
            }
        }
        free(candi_dist);
        free(candi_idx);
    }
    if (overflow) *overflow = any_overflow;
}

/* Random ball query -- random_ball_query_cuda_kernel (random_ball_query_cuda_kernel.cu:58-108):
 * first nsample hits while scanning the cloud in the caller-supplied permutation `order`. */
ORACLE_API void oracle_random_ball_query(int m, int nsample, float min_radius, float max_radius,
                                         const int *order, const float *xyz, const float *new_xyz,
                                         const int *offset, const int *new_offset, int *idx,
                                         float *dist2) {
#pragma omp parallel for schedule(static, 64)
    for (int pt = 0; pt < m; ++pt) {
        const int bt = get_bt_idx(pt, new_offset);
        const int start = bt == 0 ? 0 : offset[bt - 1];
        const int end = offset[bt];
        const float max_radius2 = max_radius * max_radius;
        const float min_radius2 = min_radius * min_radius;
        const float qx = new_xyz[pt * 3 + 0], qy = new_xyz[pt * 3 + 1], qz = new_xyz[pt * 3 + 2];
        int *o_idx = idx + (size_t)pt * nsample;
        float *o_d = dist2 + (size_t)pt * nsample;
        int cnt = 0;
        for (int i = start; i < end; i++) {
            float d2 = dist2_query(xyz, order[i], qx, qy, qz);
            if ((double)d2 <= 1e-5 || (d2 >= min_radius2 && d2 < max_radius2)) {
                o_d[cnt] = d2;
                o_idx[cnt] = order[i];
                cnt += 1;
                if (cnt >= nsample) break;
            }
        }
        for (int i = cnt; i < nsample; i++) { o_idx[i] = -1; o_d[i] = 1e10f; }
    }
}

/* ------------------------------------------------------------------------------------------
 * grouping -- grouping_cuda_kernel.cu:5-25.  forward: out[m,s,c] = in[idx[m,s],c];
 * backward: grad_in[idx[m,s],c] += grad_out[m,s,c] (atomicAdd on the GPU: order-free sum).
 * ------------------------------------------------------------------------------------------ */
ORACLE_API void oracle_grouping_forward(int m, int nsample, int c, const float *input,
                                        const int *idx, float *output) {
    for (long index = 0; index < (long)m * nsample * c; ++index) {
        const int c_idx = index % c;
        const int ns_idx = (index / c) % nsample;
        const int m_idx = index / nsample / c;
        output[index] = input[(long)idx[m_idx * nsample + ns_idx] * c + c_idx];
    }
}

ORACLE_API void oracle_grouping_backward(int m, int nsample, int c, const float *grad_output,
                                         const int *idx, float *grad_input) {
    for (long index = 0; index < (long)m * nsample * c; ++index) {
        const int c_idx = index % c;
        const int ns_idx = (index / c) % nsample;
        const int m_idx = index / nsample / c;
        grad_input[(long)idx[m_idx * nsample + ns_idx] * c + c_idx] += grad_output[index];
    }
}

/* interpolation -- interpolation_cuda_kernel.cu:5-33.  forward accumulates in k order with the
 * GPU's contracted FFMA: out = fmaf(in, w, out). */
ORACLE_API void oracle_interpolation_forward(int n, int c, int k, const float *input,
                                             const int *idx, const float *weight, float *output) {
    for (long index = 0; index < (long)n * c; ++index) {
        const int c_idx = index % c;
        const int n_idx = index / c;
        for (int i = 0; i < k; i++) {
            const int idx_idx = n_idx * k + i;
            output[index] = fmaf(input[(long)idx[idx_idx] * c + c_idx], weight[idx_idx], output[index]);
        }
    }
}

ORACLE_API void oracle_interpolation_backward(int n, int c, int k, const float *grad_output,
                                              const int *idx, const float *weight,
                                              float *grad_input) {
    for (long index = 0; index < (long)n * c; ++index) {
        const int c_idx = index % c;
        const int n_idx = index / c;
        for (int i = 0; i < k; i++) {
            const int idx_idx = n_idx * k + i;
            grad_input[(long)idx[idx_idx] * c + c_idx] += grad_output[index] * weight[idx_idx];
        }
    }
}

/* aggregation -- aggregation_cuda_kernel.cu:5-39. */
ORACLE_API void oracle_aggregation_forward(int n, int nsample, int c, int w_c, const float *input,
                                           const float *position, const float *weight,
                                           const int *idx, float *output) {
    for (long index = 0; index < (long)n * c; ++index) {
        const int c_idx = index % c;
        const int n_idx = index / c;
        const int w_c_idx = c_idx % w_c;
        for (int s = 0; s < nsample; s++) {
            const int idx_idx = n_idx * nsample + s;
            const long input_idx = (long)idx[idx_idx] * c + c_idx;
            const long position_idx = (long)n_idx * nsample * c + (long)s * c + c_idx;
            const long weight_idx = (long)n_idx * nsample * w_c + (long)s * w_c + w_c_idx;
            output[index] = fmaf(input[input_idx] + position[position_idx], weight[weight_idx], output[index]);
        }
    }
}

ORACLE_API void oracle_aggregation_backward(int n, int nsample, int c, int w_c, const float *input,
                                            const float *position, const float *weight,
                                            const int *idx, const float *grad_output,
                                            float *grad_input, float *grad_position,
                                            float *grad_weight) {
    for (long index = 0; index < (long)n * c; ++index) {
        const int c_idx = index % c;
        const int n_idx = index / c;
        const int w_c_idx = c_idx % w_c;
        for (int s = 0; s < nsample; s++) {
            const int idx_idx = n_idx * nsample + s;
            const long input_idx = (long)idx[idx_idx] * c + c_idx;
            const long position_idx = (long)n_idx * nsample * c + (long)s * c + c_idx;
            const long weight_idx = (long)n_idx * nsample * w_c + (long)s * w_c + w_c_idx;
            grad_input[input_idx] += grad_output[index] * weight[weight_idx];
            grad_position[position_idx] = grad_output[index] * weight[weight_idx];
            grad_weight[weight_idx] += grad_output[index] * (input[input_idx] + position[position_idx]);
        }
    }
}

/* subtraction -- subtraction_cuda_kernel.cu:5-30. */
ORACLE_API void oracle_subtraction_forward(int n, int nsample, int c, const float *input1,
                                           const float *input2, const int *idx, float *output) {
    for (long index = 0; index < (long)n * nsample * c; ++index) {
        const int c_idx = index % c;
        const int s = (index / c) % nsample;
        const int n_idx = index / nsample / c;
        output[index] = input1[(long)n_idx * c + c_idx] - input2[(long)idx[n_idx * nsample + s] * c + c_idx];
    }
}

ORACLE_API void oracle_subtraction_backward(int n, int nsample, int c, const int *idx,
                                            const float *grad_output, float *grad_input1,
                                            float *grad_input2) {
    for (long index = 0; index < (long)n * nsample * c; ++index) {
        const int c_idx = index % c;
        const int s = (index / c) % nsample;
        const int n_idx = index / nsample / c;
        grad_input1[(long)n_idx * c + c_idx] += grad_output[index];
        grad_input2[(long)idx[n_idx * nsample + s] * c + c_idx] += -grad_output[index];
    }
}

/* attention relation / fusion steps -- attention_cuda_kernel.cu:9-86. */
ORACLE_API void oracle_attention_relation_step_forward(int m, int g, int c, const float *query,
                                                       const float *key, const float *weight,
                                                       const int *index_target,
                                                       const int *index_refer, float *output) {
    for (int r = 0; r < m; ++r)
        for (int gi = 0; gi < g; ++gi)
            for (int ci = 0; ci < c; ++ci) {
                const long q_idx = (long)index_target[r] * g * c + gi * c + ci;
                const long k_idx = (long)index_refer[r] * g * c + gi * c + ci;
                output[(long)r * g + gi] += query[q_idx] * key[k_idx] * weight[ci];
            }
}

ORACLE_API void oracle_attention_relation_step_backward(int m, int g, int c, const float *query,
                                                        float *grad_query, const float *key,
                                                        float *grad_key, const float *weight,
                                                        float *grad_weight,
                                                        const int *index_target,
                                                        const int *index_refer,
                                                        const float *grad_output) {
    for (int r = 0; r < m; ++r)
        for (int gi = 0; gi < g; ++gi)
            for (int ci = 0; ci < c; ++ci) {
                const long q_idx = (long)index_target[r] * g * c + gi * c + ci;
                const long k_idx = (long)index_refer[r] * g * c + gi * c + ci;
                const float grad_r = grad_output[(long)r * g + gi];
                grad_query[q_idx] += grad_r * key[k_idx] * weight[ci];
                grad_key[k_idx] += grad_r * query[q_idx] * weight[ci];
                grad_weight[ci] += grad_r * key[k_idx] * query[q_idx];
            }
}

ORACLE_API void oracle_attention_fusion_step_forward(int m, int g, int c, const float *weight,
                                                     const float *value, const int *index_target,
                                                     const int *index_refer, float *output) {
    for (int r = 0; r < m; ++r)
        for (int gi = 0; gi < g; ++gi)
            for (int ci = 0; ci < c; ++ci) {
                const long o_idx = (long)index_target[r] * g * c + gi * c + ci;
                const long v_idx = (long)index_refer[r] * g * c + gi * c + ci;
                output[o_idx] += weight[(long)r * g + gi] * value[v_idx];
            }
}

ORACLE_API void oracle_attention_fusion_step_backward(int m, int g, int c, const float *weight,
                                                      float *grad_weight, const float *value,
                                                      float *grad_value, const int *index_target,
                                                      const int *index_refer,
                                                      const float *grad_output) {
    for (int r = 0; r < m; ++r)
        for (int gi = 0; gi < g; ++gi)
            for (int ci = 0; ci < c; ++ci) {
                const long o_idx = (long)index_target[r] * g * c + gi * c + ci;
                const long v_idx = (long)index_refer[r] * g * c + gi * c + ci;
                const long w_idx = (long)r * g + gi;
                const float grad = grad_output[o_idx];
                grad_weight[w_idx] += grad * value[v_idx];
                grad_value[v_idx] += grad * weight[w_idx];
            }
}
