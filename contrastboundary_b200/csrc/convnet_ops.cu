// convnet_ops.cu — the remaining device operators of the TF tree's ConvNet (config 3: AdaptiveWeight ResNet + CBL):
//   ind_max_pool          tensorflow/models/basic_operators.py:155-172   (shortcut of the strided bottleneck, resnet.py:268)
//   hard sub-scene labels  tensorflow/models/heads/head.py:25-49,117-131,518-531 (get_scene_label_infer, reduction 'max':
//                         arg-max of the label histogram of the full-resolution points around a coarse point) — either over
//                         a given neighbour matrix with shadow entries (stage 1: the pooling neighbours, get_sample_idx
//                         :154-156) or over ALL full-resolution points within a radius (stages >= 2: get_sample_idx :158-176
//                         runs a radius search with r_sample[i-1]; here the vote is taken inside the search, the
//                         (n_i x thousands) neighbour matrix is never materialised).
#include "knn.cuh"

// ---------------------------------------------------------------------------------------------
// ind_max_pool: out[i,c] = max_k x_shadow[inds[i,k], c] with x_shadow = [x ; column minima]  (shadow row = n1)
// arg[i,c] = the k that won (first maximum), saved for the backward; 255 = the shadow row won (no real neighbour).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ind_max_pool(int n2, int k, int c, int n1, const float *__restrict__ x,
                                                      const float *__restrict__ colmin, const int *__restrict__ inds,
                                                      float *__restrict__ out, unsigned char *__restrict__ arg)
{
    const int c4 = c >> 2;
    const long long total = (long long)n2 * c4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / c4), q = (int)(t % c4);
        float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned char a[4] = {255, 255, 255, 255};
        bool any = false;
        for (int kk = 0; kk < k; kk++) {
            const int j = __ldg(inds + (size_t)i * k + kk);
            float4 v;
            if (j >= 0 && j < n1) v = __ldg(reinterpret_cast<const float4 *>(x + (size_t)j * c) + q);
            else v = __ldg(reinterpret_cast<const float4 *>(colmin) + q);
            const unsigned char id = (j >= 0 && j < n1) ? (unsigned char)kk : (unsigned char)255;
            if (!any) { best = v; a[0] = a[1] = a[2] = a[3] = id; any = true; continue; }
            if (v.x > best.x) { best.x = v.x; a[0] = id; }
            if (v.y > best.y) { best.y = v.y; a[1] = id; }
            if (v.z > best.z) { best.z = v.z; a[2] = id; }
            if (v.w > best.w) { best.w = v.w; a[3] = id; }
        }
        reinterpret_cast<float4 *>(out + (size_t)i * c)[q] = best;
        reinterpret_cast<uchar4 *>(arg + (size_t)i * c)[q] = make_uchar4(a[0], a[1], a[2], a[3]);
    }
}

// backward as a GATHER over the inverse map would need the inverse neighbour list; the pooled sets of distinct coarse
// points overlap (a fine point lies within the pooling radius of several coarse points), so this is a scatter-add.
__global__ void __launch_bounds__(256) k_ind_max_pool_bwd(int n2, int k, int c, const int *__restrict__ inds,
                                                          const unsigned char *__restrict__ arg, const float *__restrict__ gout,
                                                          float *__restrict__ gx)
{
    const long long total = (long long)n2 * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / c), ch = (int)(t % c);
        const unsigned char a = arg[t];
        if (a == 255) continue;                      // the shadow row (column minimum) won: no real neighbour in this row
        const int j = __ldg(inds + (size_t)i * k + a);
        atomicAdd(gx + (size_t)j * c + ch, gout[t]);
    }
}

extern "C" int cb_ind_max_pool_forward(int n2, int k, int c, int n1, const float *x, const float *colmin, const int *inds,
                                       float *out, unsigned char *arg, void *stream)
{
    CB_REQUIRE(n2 >= 0 && k > 0 && k < 255 && c > 0 && c % 4 == 0 && x && colmin && inds && out && arg, CB_EINVAL,
               "cb_ind_max_pool_forward: bad arguments (c % 4 == 0, k < 255)");
    if (n2 == 0) return CB_OK;
    long long g = ((long long)n2 * (c / 4) + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_ind_max_pool<<<(int)g, 256, 0, (cudaStream_t)stream>>>(n2, k, c, n1, x, colmin, inds, out, arg);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_ind_max_pool_forward");
    return CB_OK;
}

extern "C" int cb_ind_max_pool_backward(int n2, int k, int c, const int *inds, const unsigned char *arg, const float *grad_out,
                                        float *grad_x, void *stream)
{
    CB_REQUIRE(n2 >= 0 && k > 0 && c > 0 && inds && arg && grad_out && grad_x, CB_EINVAL, "cb_ind_max_pool_backward: bad arguments");
    if (n2 == 0) return CB_OK;
    long long g = ((long long)n2 * c + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_ind_max_pool_bwd<<<(int)g, 256, 0, (cudaStream_t)stream>>>(n2, k, c, inds, arg, grad_out, grad_x);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_ind_max_pool_backward");
    return CB_OK;
}

// ---------------------------------------------------------------------------------------------
// hard sub-scene labels
// ---------------------------------------------------------------------------------------------
// over a neighbour matrix with shadow entries (idx >= n_valid or < 0 are skipped: tf_gather(..., shadow_fn=-1) then
// one_hot(-1) = 0, head.py:38-40); a row without any valid neighbour gets class 0 (arg-max of an all-zero histogram)
__global__ void k_label_vote_idx(int m, int kr, int ncls, int n_valid, const int *__restrict__ label_idx,
                                 const long long *__restrict__ target, int *__restrict__ cls)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int cnt[64];
    for (int c = 0; c < ncls; c++) cnt[c] = 0;
    for (int t = 0; t < kr; t++) {
        const int j = __ldg(label_idx + (size_t)i * kr + t);
        if (j < 0 || j >= n_valid) continue;
        const int l = (int)target[j];
        if (l >= 0 && l < ncls) cnt[l]++;
    }
    int best = 0;
    for (int c = 1; c < ncls; c++) if (cnt[c] > cnt[best]) best = c;
    cls[i] = best;
}

extern "C" int cb_label_vote_idx(int m, int kr, int ncls, int n_valid, const int *label_idx, const long long *target, int *cls,
                                 void *stream)
{
    CB_REQUIRE(m >= 0 && kr > 0 && ncls > 0 && ncls <= 64 && label_idx && target && cls, CB_EINVAL, "cb_label_vote_idx: bad arguments");
    if (m == 0) return CB_OK;
    k_label_vote_idx<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(m, kr, ncls, n_valid, label_idx, target, cls);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_label_vote_idx");
    return CB_OK;
}

// over every support within the radius (strict d2 < r^2 in nanoflann's arithmetic, as cb_radius_count): warp per query,
// lanes stride over the candidates of the covered cell rows and vote into a per-warp shared histogram
__global__ void __launch_bounds__(128) k_label_vote_radius(int m, const float *__restrict__ q_xyz, const int *__restrict__ q_offset,
                                                           int b, const CbScene *__restrict__ scenes, const int *__restrict__ cells,
                                                           const float4 *__restrict__ sorted, float radius, float r2, int ncls,
                                                           const long long *__restrict__ target, int *__restrict__ cls)
{
    __shared__ int hist[4][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * 4 + wib;
    if (w >= m) return;
    for (int c = lane; c < ncls; c += 32) hist[wib][c] = 0;
    __syncwarp();
    const float qx = __ldg(q_xyz + 3 * w), qy = __ldg(q_xyz + 3 * w + 1), qz = __ldg(q_xyz + 3 * w + 2);
    const CbScene sc = scenes[cb_scene_of(w, q_offset, b)];
    const float fx = cb_cellf(qx, sc.ox, sc.inv_h), fy = cb_cellf(qy, sc.oy, sc.inv_h), fz = cb_cellf(qz, sc.oz, sc.inv_h);
    const float rc = radius * sc.inv_h + 0.05f;          // radius in cells (+ margin for fp32 cell assignment)
    const int x0 = max((int)floorf(fx - rc), 0), x1 = min((int)floorf(fx + rc), sc.nx - 1);
    const int y0 = max((int)floorf(fy - rc), 0), y1 = min((int)floorf(fy + rc), sc.ny - 1);
    const int z0 = max((int)floorf(fz - rc), 0), z1 = min((int)floorf(fz + rc), sc.nz - 1);
    if (x0 <= x1)
        for (int z = z0; z <= z1; z++)
            for (int y = y0; y <= y1; y++) {
                const int rowbase = sc.cell_base + (z * sc.ny + y) * sc.nx;
                const int s = __ldg(cells + rowbase + x0), e = __ldg(cells + rowbase + x1 + 1);
                for (int i = s + lane; i < e; i += 32) {
                    const float4 p = __ldg(sorted + i);
                    if (cb_sqdist_mode(1, qx, qy, qz, p.x, p.y, p.z) < r2) {
                        const int l = (int)target[__float_as_int(p.w)];
                        if (l >= 0 && l < ncls) atomicAdd(&hist[wib][l], 1);
                    }
                }
            }
    __syncwarp();
    if (lane == 0) {
        int best = 0;
        for (int c = 1; c < ncls; c++) if (hist[wib][c] > hist[wib][best]) best = c;
        cls[w] = best;
    }
}

// queries (nq,3) / supports (ns,3) with cumulative scene ends; target (ns) int64 labels of the supports;
// workspace >= cb_knn_workspace_bytes(ns, 0, b) (the support grid is built here)
extern "C" int cb_label_vote_radius(int nq, const float *queries, int ns, const float *supports, const int *q_offset,
                                    const int *s_offset, int b, float radius, int ncls, const long long *target, int *cls,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    CB_REQUIRE(nq >= 0 && ns >= 0 && b > 0 && ncls > 0 && ncls <= 64 && q_offset && s_offset && target && cls && workspace && radius > 0.f,
               CB_EINVAL, "cb_label_vote_radius: bad arguments");
    CB_REQUIRE(((uintptr_t)workspace & 255) == 0, CB_EINVAL, "cb_label_vote_radius: workspace not 256-byte aligned");
    CbGridView v;
    const size_t need = cb_grid_layout(ns, 0, b, workspace, &v);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_label_vote_radius: workspace %zu < %zu", workspace_bytes, need);
    if (nq == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cb_grid_build_impl(supports, ns, s_offset, b, 32, v, st);
    if (rc) return rc;
    k_label_vote_radius<<<(nq + 3) / 4, 128, 0, st>>>(nq, queries, q_offset, b, v.scenes, v.cells, v.sorted, radius,
                                                      radius * radius, ncls, target, cls);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_label_vote_radius");
    return CB_OK;
}
