/*
 * oracle/tfops_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's TensorFlow-side CPU operators
 * (LiyaoTang/contrastBoundary, tensorflow/ops/...):
 *   - grid_subsampling, CPython flavour  (cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106)
 *   - batch_grid_subsampling, TF flavour (tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp:6-162)
 *   - batch_nanoflann_neighbors          (tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:213-336,
 *                                         distance = nanoflann.hpp:432-440, test = nanoflann.hpp:249-251)
 *   - knn_batch (nanoflann KNN)          (nearest_neighbors/knn_.cxx:72-135)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  It is a restatement (brute force instead of a KD-tree, own data structures), pinned
 * against the reference's own sources compiled unmodified (oracle/_ref/libref_cpu.so, see
 * oracle/build_ref.sh) by tests/test_oracle_vs_ref.py and the vectors under tests/golden/.
 *
 * Ordering note: the reference emits subsampled points in the iteration order of a libstdc++
 * std::unordered_map<size_t, ...>; we reproduce that by using the same container for the key
 * set (same libstdc++ → same order); everything else is our own code.
 * Build with -ffp-contract=off: the reference is built -O2 without -mfma (compile_op.sh:26-31),
 * so its float arithmetic has no fused multiply-adds.
 */
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <unordered_map>
#include <vector>

namespace {

struct Cell {
    int count = 0;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    std::vector<float> feat;
    std::vector<std::unordered_map<int, int>> labels;   // per label column histogram
};

struct GridSpec {
    float ox, oy, oz;
    size_t nx, ny;
};

// grid_subsampling.cpp:25-32 — origin aligned to the cell size, NX / NY from the bbox
GridSpec grid_spec(const float *pts, size_t n, float dl)
{
    float mnx = pts[0], mny = pts[1], mnz = pts[2], mxx = pts[0], mxy = pts[1];
    for (size_t i = 0; i < n; i++) {
        const float *p = pts + 3 * i;
        if (p[0] < mnx) mnx = p[0];
        if (p[1] < mny) mny = p[1];
        if (p[2] < mnz) mnz = p[2];
        if (p[0] > mxx) mxx = p[0];
        if (p[1] > mxy) mxy = p[1];
    }
    const float inv = 1 / dl;                       // (1/sampleDl): int / float → float
    GridSpec g;
    g.ox = std::floor(mnx * inv) * dl;
    g.oy = std::floor(mny * inv) * dl;
    g.oz = std::floor(mnz * inv) * dl;
    g.nx = (size_t)std::floor((mxx - g.ox) / dl) + 1;
    g.ny = (size_t)std::floor((mxy - g.oy) / dl) + 1;
    return g;
}

inline size_t cell_key(const GridSpec &g, const float *p, float dl)
{
    size_t ix = (size_t)std::floor((p[0] - g.ox) / dl);
    size_t iy = (size_t)std::floor((p[1] - g.oy) / dl);
    size_t iz = (size_t)std::floor((p[2] - g.oz) / dl);
    return ix + g.nx * iy + g.nx * g.ny * iz;
}

// One scene.  Returns number of occupied cells; appends to the output vectors.
size_t subsample_scene(const float *pts, size_t n, const float *feat, size_t fdim, const int *cls,
                       size_t ldim, float dl, std::vector<float> &out_pts, std::vector<float> &out_feat,
                       std::vector<int> &out_cls)
{
    if (n == 0) return 0;
    GridSpec g = grid_spec(pts, n, dl);
    std::unordered_map<size_t, Cell> data;          // same container/key type as the reference → same order
    for (size_t i = 0; i < n; i++) {
        const float *p = pts + 3 * i;
        size_t key = cell_key(g, p, dl);
        auto it = data.find(key);
        if (it == data.end()) {
            Cell c;
            c.feat.assign(fdim, 0.f);
            c.labels.resize(ldim);
            it = data.emplace(key, std::move(c)).first;
        }
        Cell &c = it->second;
        c.count += 1;
        c.sx += p[0]; c.sy += p[1]; c.sz += p[2];
        for (size_t f = 0; f < fdim; f++) c.feat[f] += feat[i * fdim + f];
        for (size_t l = 0; l < ldim; l++) c.labels[l][cls[i * ldim + l]] += 1;
    }
    for (auto &kv : data) {
        Cell &c = kv.second;
        const float a = (float)(1.0 / c.count);     // double 1.0/count narrowed by operator*(PointXYZ, float)
        out_pts.push_back(c.sx * a);
        out_pts.push_back(c.sy * a);
        out_pts.push_back(c.sz * a);
        const float cnt = (float)c.count;
        for (size_t f = 0; f < fdim; f++) out_feat.push_back(c.feat[f] / cnt);
        for (size_t l = 0; l < ldim; l++) {
            // std::max_element with (a.second < b.second): first maximum in the histogram map's order
            auto best = c.labels[l].begin();
            for (auto h = c.labels[l].begin(); h != c.labels[l].end(); ++h)
                if (best->second < h->second) best = h;
            out_cls.push_back(best->first);
        }
    }
    return data.size();
}

}  // namespace

extern "C" {

/*
 * CPython-flavour grid_subsampling (features mean + per-column label majority vote).
 * Outputs must have room for n rows; returns the number of rows written.
 * feat/cls may be NULL (fdim/ldim = 0).
 */
int oracle_grid_subsample(const float *points, int n, const float *features, int fdim, const int *classes,
                          int ldim, float dl, float *out_points, float *out_features, int *out_classes)
{
    std::vector<float> op, of;
    std::vector<int> oc;
    size_t m = subsample_scene(points, (size_t)n, features, features ? (size_t)fdim : 0, classes,
                               classes ? (size_t)ldim : 0, dl, op, of, oc);
    std::memcpy(out_points, op.data(), op.size() * sizeof(float));
    if (features && out_features) std::memcpy(out_features, of.data(), of.size() * sizeof(float));
    if (classes && out_classes) std::memcpy(out_classes, oc.data(), oc.size() * sizeof(int));
    return (int)m;
}

/* TF-flavour batch_grid_subsampling: `batches` are per-scene LENGTHS (not cumulative offsets). */
int oracle_batch_grid_subsample(const float *points, int n, const int *batches, int b, float dl,
                                float *out_points, int *out_batches)
{
    (void)n;
    std::vector<float> op, of;
    std::vector<int> oc;
    size_t start = 0;
    for (int i = 0; i < b; i++) {
        size_t m = subsample_scene(points + 3 * start, (size_t)batches[i], nullptr, 0, nullptr, 0, dl, op, of, oc);
        out_batches[i] = (int)m;
        start += (size_t)batches[i];
    }
    std::memcpy(out_points, op.data(), op.size() * sizeof(float));
    return (int)(op.size() / 3);
}

/*
 * batch radius neighbours.  Returns a malloc'd (nq, *max_count) int32 matrix (free with
 * oracle_free); rows sorted by ascending squared distance, ties broken by ascending index
 * (the reference's std::sort leaves tie order unspecified), padded with ns.
 */
int *oracle_batch_radius_neighbors(const float *queries, int nq, const float *supports, int ns,
                                   const int *q_batches, const int *s_batches, int b, float radius,
                                   int *max_count_out)
{
    const float r2 = radius * radius;
    std::vector<std::vector<std::pair<float, int>>> rows((size_t)nq);
    std::vector<int> qstart(b + 1, 0), sstart(b + 1, 0);
    for (int i = 0; i < b; i++) { qstart[i + 1] = qstart[i] + q_batches[i]; sstart[i + 1] = sstart[i] + s_batches[i]; }
    for (int bi = 0; bi < b; bi++) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int q = qstart[bi]; q < qstart[bi + 1]; q++) {
            const float *a = queries + 3 * (size_t)q;
            auto &row = rows[(size_t)q];
            for (int s = sstart[bi]; s < sstart[bi + 1]; s++) {
                const float *p = supports + 3 * (size_t)s;
                float d = 0.f;
                float d0 = a[0] - p[0]; d += d0 * d0;
                float d1 = a[1] - p[1]; d += d1 * d1;
                float d2 = a[2] - p[2]; d += d2 * d2;
                if (d < r2) row.emplace_back(d, s);
            }
            std::sort(row.begin(), row.end());
        }
    }
    size_t mc = 0;
    for (auto &r : rows) mc = std::max(mc, r.size());
    *max_count_out = (int)mc;
    int *out = (int *)std::malloc(sizeof(int) * std::max<size_t>((size_t)nq * mc, 1));
    for (int q = 0; q < nq; q++)
        for (size_t j = 0; j < mc; j++)
            out[(size_t)q * mc + j] = j < rows[(size_t)q].size() ? rows[(size_t)q][j].second : ns;
    return out;
}

/*
 * knn_batch (nearest_neighbors/knn_.cxx:72-135): fixed-size batches [B,N,3] supports, [B,M,3]
 * queries → [B,M,K] indices local to the batch element, ascending distance.
 */
void oracle_knn_batch(const float *supports, const float *queries, int B, int N, int M, int K, long long *out)
{
#pragma omp parallel for schedule(dynamic, 16) collapse(2)
    for (int bi = 0; bi < B; bi++)
        for (int q = 0; q < M; q++) {
            const float *a = queries + 3 * ((size_t)bi * M + q);
            std::vector<std::pair<float, int>> row((size_t)N);
            for (int s = 0; s < N; s++) {
                const float *p = supports + 3 * ((size_t)bi * N + s);
                float d = 0.f;
                float d0 = a[0] - p[0]; d += d0 * d0;
                float d1 = a[1] - p[1]; d += d1 * d1;
                float d2 = a[2] - p[2]; d += d2 * d2;
                row[(size_t)s] = {d, s};
            }
            int kk = std::min(K, N);
            std::partial_sort(row.begin(), row.begin() + kk, row.end());
            for (int j = 0; j < K; j++) out[((size_t)bi * M + q) * K + j] = j < kk ? row[(size_t)j].second : 0;
        }
}

void oracle_free(void *p) { std::free(p); }

}  // extern "C"
