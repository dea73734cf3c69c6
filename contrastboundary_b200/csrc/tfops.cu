// tfops.cu — a11/a12: device versions of the reference's TF-side CPU operators
//   batch_grid_subsampling   tensorflow/ops/tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp:6-162
//   grid_subsampling (CPython flavour: + feature mean)   tensorflow/ops/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106
//   batch_ordered_neighbors (radius search)              tensorflow/ops/tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:213-336
// The reference runs these single-threaded on the host inside tf.data; here they are CUDA kernels over
// device arrays.  Bit-exact barycentres need the reference's fp32 summation ORDER (arrival order inside a
// voxel), so points are stably sorted by voxel key and each voxel is summed sequentially by one thread.
// The reference's OUTPUT ORDER is the iteration order of a libstdc++ unordered_map; cb_unordered_map_order
// replays the key insertions into the same container on the host (O(#voxels)) to obtain that permutation.
#include "knn.cuh"
#include <cub/cub.cuh>
#include <unordered_map>
#include <vector>
#include <algorithm>

// ---------------------------------------------------------------------------------------------
// radius neighbours
// ---------------------------------------------------------------------------------------------
// count supports with d2 < r2 per query (warp per query), and the global maximum
__global__ void __launch_bounds__(128) k_radius_count(int m, const float *__restrict__ q_xyz, const int *__restrict__ q_offset,
                                                      int b, const CbScene *__restrict__ scenes,
                                                      const int *__restrict__ cells, const float4 *__restrict__ sorted,
                                                      float radius, float r2, int *__restrict__ counts, int *__restrict__ max_count)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= m) return;
    const float qx = __ldg(q_xyz + 3 * w), qy = __ldg(q_xyz + 3 * w + 1), qz = __ldg(q_xyz + 3 * w + 2);
    const CbScene sc = scenes[cb_scene_of(w, q_offset, b)];
    const float fx = cb_cellf(qx, sc.ox, sc.inv_h), fy = cb_cellf(qy, sc.oy, sc.inv_h), fz = cb_cellf(qz, sc.oz, sc.inv_h);
    const float rc = radius * sc.inv_h + 0.05f;          // radius in cells (+ margin for fp32 cell assignment)
    const int x0 = max((int)floorf(fx - rc), 0), x1 = min((int)floorf(fx + rc), sc.nx - 1);
    const int y0 = max((int)floorf(fy - rc), 0), y1 = min((int)floorf(fy + rc), sc.ny - 1);
    const int z0 = max((int)floorf(fz - rc), 0), z1 = min((int)floorf(fz + rc), sc.nz - 1);
    int cnt = 0;
    if (x0 <= x1)
        for (int z = z0; z <= z1; z++)
            for (int y = y0; y <= y1; y++) {
                const int rowbase = sc.cell_base + (z * sc.ny + y) * sc.nx;
                const int s = __ldg(cells + rowbase + x0), e = __ldg(cells + rowbase + x1 + 1);
                for (int i = s + lane; i < e; i += 32) {
                    const float4 p = __ldg(sorted + i);
                    cnt += cb_sqdist_mode(1, qx, qy, qz, p.x, p.y, p.z) < r2;
                }
            }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(CB_FULL_MASK, cnt, o);
    if (lane == 0) {
        counts[w] = cnt;
        atomicMax(max_count, cnt);
    }
}

extern "C" int cb_radius_count(int nq, const float *queries, int ns, const float *supports, const int *q_offset,
                               const int *s_offset, int b, float radius, int *counts, int *max_count, void *workspace,
                               size_t workspace_bytes, void *stream)
{
    CB_REQUIRE(nq >= 0 && ns >= 0 && b > 0 && q_offset && s_offset && counts && max_count && workspace, CB_EINVAL,
               "cb_radius_count: bad arguments");
    CbGridView v;
    const size_t need = cb_grid_layout(ns, nq, b, workspace, &v);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_radius_count: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(max_count, 0, sizeof(int), st);
    int rc = cb_grid_build_impl(supports, ns, s_offset, b, 24, v, st);
    if (rc) return rc;
    if (nq > 0)
        k_radius_count<<<(nq + 3) / 4, 128, 0, st>>>(nq, queries, q_offset, b, v.scenes, v.cells, v.sorted, radius,
                                                      radius * radius, counts, max_count);
    CB_COUNT(2);
    CB_CUDA_CHECK("cb_radius_count");
    return CB_OK;
}

// rows of `width` nearest supports with d2 < r2 (ascending), padded with ns; grid must have been built by
// cb_radius_count on the same workspace
extern "C" int cb_radius_fill(int nq, int width, const float *queries, int ns, const float *supports, const int *q_offset,
                              const int *s_offset, int b, float radius, int *neighbors, void *workspace,
                              size_t workspace_bytes, void *stream)
{
    CB_REQUIRE(nq >= 0 && width >= 0 && neighbors && workspace, CB_EINVAL, "cb_radius_fill: bad arguments");
    CbGridView v;
    const size_t need = cb_grid_layout(ns, nq, b, workspace, &v);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_radius_fill: workspace %zu < %zu", workspace_bytes, need);
    return cb_knn_query_radius_impl(nq, width, supports, ns, queries, s_offset, q_offset, b, neighbors, radius * radius, ns, v,
                                    (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// grid subsampling
// ---------------------------------------------------------------------------------------------
struct GsScene {
    float ox, oy, oz;
    unsigned long long nx, ny;
    int start, end;
};

// originCorner = floor(minCorner * (1/dl)) * dl ; NX = floor((max.x - origin.x)/dl) + 1  (grid_subsampling.cpp:25-32)
__global__ void k_gs_scenes(const unsigned *__restrict__ bbox, const int *__restrict__ offset, int b, float dl, GsScene *sc)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b) return;
    GsScene g;
    g.start = s == 0 ? 0 : offset[s - 1];
    g.end = offset[s];
    if (g.end <= g.start) { g.ox = g.oy = g.oz = 0.f; g.nx = g.ny = 1; sc[s] = g; return; }
    const float inv = __fdiv_rn(1.0f, dl);
    const float mnx = cb_ord2f(bbox[6 * s]), mny = cb_ord2f(bbox[6 * s + 1]), mnz = cb_ord2f(bbox[6 * s + 2]);
    const float mxx = cb_ord2f(bbox[6 * s + 3]), mxy = cb_ord2f(bbox[6 * s + 4]);
    g.ox = __fmul_rn(floorf(__fmul_rn(mnx, inv)), dl);
    g.oy = __fmul_rn(floorf(__fmul_rn(mny, inv)), dl);
    g.oz = __fmul_rn(floorf(__fmul_rn(mnz, inv)), dl);
    g.nx = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mxx, g.ox), dl)) + 1ull;
    g.ny = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mxy, g.oy), dl)) + 1ull;
    sc[s] = g;
}

// sort key = scene << 44 | voxel key  (voxel key must be < 2^44)
__global__ void k_gs_keys(const float *__restrict__ xyz, int n, const int *__restrict__ offset, int b, const GsScene *__restrict__ sc,
                          float dl, unsigned long long *keys, int *vals, int *overflow)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = cb_scene_of(i, offset, b);
    const GsScene g = sc[s];
    const unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i], g.ox), dl));
    const unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i + 1], g.oy), dl));
    const unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i + 2], g.oz), dl));
    const unsigned long long key = ix + g.nx * iy + g.nx * g.ny * iz;       // grid_subsampling.cpp:61-64
    if (key >> 44) atomicExch(overflow, 1);
    keys[i] = ((unsigned long long)s << 44) | (key & ((1ull << 44) - 1));
    vals[i] = i;
}

__global__ void k_gs_heads(const unsigned long long *__restrict__ keys, int n, int *flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// cell c = [cell_start[c], cell_start[c+1]) in sorted order; one thread per cell sums in arrival order
__global__ void k_gs_reduce(const float *__restrict__ xyz, const float *__restrict__ feat, int fdim, int n, int ncells,
                            const int *__restrict__ flags_scan /* exclusive scan of heads */, const int *__restrict__ heads,
                            const int *__restrict__ vals, const unsigned long long *__restrict__ keys,
                            int *__restrict__ cell_first /* sorted position of the cell head */, int *__restrict__ point_cell)
{
    // record the head position of every cell and the cell (key order) of every input point
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (heads[i]) cell_first[flags_scan[i]] = i;
    if (point_cell) point_cell[vals[i]] = flags_scan[i] - (heads[i] ? 0 : 1);
    (void)xyz; (void)feat; (void)fdim; (void)ncells; (void)keys;
}

__global__ void k_gs_permute(int ncells, int fdim, const int *__restrict__ perm, const float *__restrict__ in_xyz,
                             const float *__restrict__ in_feat, float *__restrict__ out_xyz, float *__restrict__ out_feat)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int src = perm[c];
    out_xyz[3 * c] = in_xyz[3 * src]; out_xyz[3 * c + 1] = in_xyz[3 * src + 1]; out_xyz[3 * c + 2] = in_xyz[3 * src + 2];
    if (in_feat)
        for (int f = 0; f < fdim; f++) out_feat[(size_t)c * fdim + f] = in_feat[(size_t)src * fdim + f];
}


__global__ void k_gs_ncells(const int *scan, const int *flags, int n, int *out) { out[0] = scan[n - 1] + flags[n - 1]; }

__global__ void k_gs_bary_dyn(const float *xyz, const float *feat, int fdim, int n, const int *ncells_dev, const int *cell_first,
                              const int *vals, const unsigned long long *keys, float *out_xyz, float *out_feat,
                              unsigned long long *cell_key, int *cell_first_idx, int *cell_count)
{
    const int ncells = *ncells_dev;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int s = cell_first[c], e = c + 1 < ncells ? cell_first[c + 1] : n;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int t = s; t < e; t++) {
        const int i = vals[t];
        sx = __fadd_rn(sx, xyz[3 * i]); sy = __fadd_rn(sy, xyz[3 * i + 1]); sz = __fadd_rn(sz, xyz[3 * i + 2]);
    }
    const int cnt = e - s;
    const float a = (float)(1.0 / (double)cnt);
    out_xyz[3 * c] = __fmul_rn(sx, a); out_xyz[3 * c + 1] = __fmul_rn(sy, a); out_xyz[3 * c + 2] = __fmul_rn(sz, a);
    if (feat) {
        const float fc = (float)cnt;
        for (int f = 0; f < fdim; f++) {
            float acc = 0.f;
            for (int t = s; t < e; t++) acc = __fadd_rn(acc, feat[(size_t)vals[t] * fdim + f]);
            out_feat[(size_t)c * fdim + f] = __fdiv_rn(acc, fc);
        }
    }
    cell_key[c] = keys[s];
    cell_first_idx[c] = vals[s];
    cell_count[c] = cnt;
}

static size_t gs_layout(int n, int b, char *base, GsScene **sc, unsigned **bbox, int **occ, CbGridHeader **hdr,
                        unsigned long long **keys, unsigned long long **keys2, int **vals, int **vals2, int **flags,
                        int **scan, int **cell_first, unsigned long long **cell_key, int **cell_first_idx, int **cell_count,
                        float **tmp_xyz, int **misc, void **cub_tmp, size_t *cub_bytes)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = cb_align_up(off + bytes, 256); return base ? (void *)(base + o) : (void *)nullptr; };
    *sc = (GsScene *)take(sizeof(GsScene) * (size_t)b);
    *bbox = (unsigned *)take(sizeof(unsigned) * 6 * (size_t)b);
    *occ = (int *)take(sizeof(int) * 2 * (size_t)b);
    *hdr = (CbGridHeader *)take(sizeof(CbGridHeader));
    *keys = (unsigned long long *)take(8 * (size_t)n);
    *keys2 = (unsigned long long *)take(8 * (size_t)n);
    *vals = (int *)take(4 * (size_t)n);
    *vals2 = (int *)take(4 * (size_t)n);
    *flags = (int *)take(4 * (size_t)n);
    *scan = (int *)take(4 * ((size_t)n + 1));
    *cell_first = (int *)take(4 * ((size_t)n + 1));
    *cell_key = (unsigned long long *)take(8 * (size_t)n);
    *cell_first_idx = (int *)take(4 * (size_t)n);
    *cell_count = (int *)take(4 * (size_t)n);
    *tmp_xyz = (float *)take(12 * (size_t)n);
    *misc = (int *)take(64);
    size_t s1 = 0, s2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, s1, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr,
                                    (int *)nullptr, n);
    cub::DeviceScan::ExclusiveSum(nullptr, s2, (int *)nullptr, (int *)nullptr, n);
    *cub_bytes = s1 > s2 ? s1 : s2;
    *cub_tmp = take(*cub_bytes + 256);
    return off;
}

extern "C" size_t cb_grid_subsample_workspace_bytes(int n, int b, int fdim)
{
    GsScene *sc; unsigned *bbox; int *occ; CbGridHeader *hdr; unsigned long long *k1, *k2, *ck; int *v1, *v2, *fl, *scn, *cf, *cfi, *cc, *misc;
    float *tx; void *ct; size_t cb;
    size_t base = gs_layout(n, b, nullptr, &sc, &bbox, &occ, &hdr, &k1, &k2, &v1, &v2, &fl, &scn, &cf, &ck, &cfi, &cc, &tx, &misc, &ct, &cb);
    return base + cb_align_up((size_t)n * (size_t)(fdim > 0 ? fdim : 1) * 4, 256) + 256;
}

// Step 1 (device, async): voxel keys, stable sort, per-voxel barycentres (+ feature means) in KEY order.
// Writes ncells to ncells_dev; cell_key / cell_first_idx / cell_scene-sorted arrays stay in the workspace.
extern "C" int cb_grid_subsample_cells(const float *xyz, int n, const int *offset, int b, float dl, const float *feat, int fdim,
                                       float *cells_xyz, float *cells_feat, unsigned long long *cells_key,
                                       int *cells_first_idx, int *point_cell, int *ncells_dev, void *workspace,
                                       size_t workspace_bytes, void *stream)
{
    CB_REQUIRE(n >= 0 && b > 0 && dl > 0.f && offset && cells_xyz && cells_key && cells_first_idx && ncells_dev && workspace,
               CB_EINVAL, "cb_grid_subsample_cells: bad arguments");
    CB_REQUIRE(b < (1 << 19), CB_EUNSUPPORTED, "cb_grid_subsample_cells: too many scenes");
    cudaStream_t st = (cudaStream_t)stream;
    GsScene *sc; unsigned *bbox; int *occ; CbGridHeader *hdr; unsigned long long *keys, *keys2, *ck; int *vals, *vals2, *flags, *scan, *cf, *cfi, *cc, *misc;
    float *tx; void *ct; size_t cb;
    const size_t need = gs_layout(n, b, (char *)workspace, &sc, &bbox, &occ, &hdr, &keys, &keys2, &vals, &vals2, &flags, &scan, &cf,
                                  &ck, &cfi, &cc, &tx, &misc, &ct, &cb);
    CB_REQUIRE(workspace_bytes >= need, CB_EWORKSPACE, "cb_grid_subsample_cells: workspace %zu < %zu", workspace_bytes, need);
    cudaMemsetAsync(misc, 0, 64, st);
    if (n == 0) { cudaMemsetAsync(ncells_dev, 0, sizeof(int), st); return CB_OK; }
    k_bbox_init<<<(b * 6 + 127) / 128, 128, 0, st>>>(bbox, occ, b, hdr, n);
    int g = (n + 255) / 256;
    k_bbox<<<g > 148 * 16 ? 148 * 16 : g, 256, 0, st>>>(xyz, n, offset, b, bbox);
    k_gs_scenes<<<(b + 127) / 128, 128, 0, st>>>(bbox, offset, b, dl, sc);
    k_gs_keys<<<g, 256, 0, st>>>(xyz, n, offset, b, sc, dl, keys, vals, misc);
    size_t tb = cb + 256;
    cub::DeviceRadixSort::SortPairs(ct, tb, keys, keys2, vals, vals2, n, 0, 64, st);     // stable
    k_gs_heads<<<g, 256, 0, st>>>(keys2, n, flags);
    tb = cb + 256;
    cub::DeviceScan::ExclusiveSum(ct, tb, flags, scan, n, st);
    k_gs_reduce<<<g, 256, 0, st>>>(xyz, feat, fdim, n, 0, scan, flags, vals2, keys2, cf, point_cell);
    // ncells = scan[n-1] + flags[n-1] stays on the device; the barycentre kernel is launched over n threads
    k_gs_ncells<<<1, 1, 0, st>>>(scan, flags, n, ncells_dev);
    k_gs_bary_dyn<<<g, 256, 0, st>>>(xyz, feat, fdim, n, ncells_dev, cf, vals2, keys2, cells_xyz, cells_feat, cells_key,
                                     cells_first_idx, cc);
    CB_COUNT(9);
    CB_CUDA_CHECK("cb_grid_subsample_cells");
    return CB_OK;
}

// Step 2 (host): the reference's output order.  keys[c] = scene << 44 | voxel key, first_idx[c] = index of the
// first input point of the voxel.  Per scene, the voxels are inserted in order of first appearance into a
// std::unordered_map<size_t, int> (exactly what the reference's loop does) and read back by iterating it.
// perm[out_pos] = cell index; scene_counts[s] = number of voxels of scene s.
extern "C" int cb_unordered_map_order(const unsigned long long *keys, const int *first_idx, int ncells, int b, int *perm,
                                      int *scene_counts)
{
    CB_REQUIRE(ncells >= 0 && b > 0 && perm && scene_counts && (ncells == 0 || (keys && first_idx)), CB_EINVAL,
               "cb_unordered_map_order: bad arguments");
    std::vector<std::vector<int>> per_scene((size_t)b);
    for (int c = 0; c < ncells; c++) per_scene[(size_t)(keys[c] >> 44)].push_back(c);
    int out = 0;
    for (int s = 0; s < b; s++) {
        std::vector<int> &cells = per_scene[(size_t)s];
        std::sort(cells.begin(), cells.end(), [&](int a, int c2) { return first_idx[a] < first_idx[c2]; });
        std::unordered_map<size_t, int> m;
        for (int c : cells) m.emplace((size_t)(keys[c] & ((1ull << 44) - 1)), c);
        for (auto &kv : m) perm[out++] = kv.second;
        scene_counts[s] = (int)cells.size();
    }
    return CB_OK;
}

extern "C" int cb_grid_subsample_permute(int ncells, int fdim, const int *perm, const float *in_xyz, const float *in_feat,
                                         float *out_xyz, float *out_feat, void *stream)
{
    if (ncells <= 0) return CB_OK;
    CB_REQUIRE(perm && in_xyz && out_xyz, CB_EINVAL, "cb_grid_subsample_permute: NULL pointer");
    k_gs_permute<<<(ncells + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ncells, fdim, perm, in_xyz, in_feat, out_xyz, out_feat);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_grid_subsample_permute");
    return CB_OK;
}

// first maximum of a label histogram in the iteration order of a std::unordered_map<int,int> filled in arrival
// order (what the reference's max_element over SampledData::labels sees; grid_subsampling.cpp:97-102).  HOST.
extern "C" int cb_label_vote_host(const int *labels, int count)
{
    std::unordered_map<int, int> h;
    for (int i = 0; i < count; i++) h[labels[i]] += 1;
    auto best = h.begin();
    for (auto it = h.begin(); it != h.end(); ++it)
        if (best->second < it->second) best = it;
    return best == h.end() ? 0 : best->first;
}
