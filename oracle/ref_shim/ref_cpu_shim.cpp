/*
 * oracle/ref_shim/ref_cpu_shim.cpp — TEST INFRASTRUCTURE.
 * extern "C" entry points over the REFERENCE's own TF-side C++ cores, which are compiled
 * unmodified from /root/reference by oracle/build_ref.sh into oracle/_ref/libref_cpu.so.
 * (The reference's TF / CPython wrappers cannot be built here: no TensorFlow, NumPy 2.)
 * Wraps:  batch_grid_subsampling   tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.h
 *         batch_nanoflann_neighbors tf_custom_ops/tf_neighbors/neighbors/neighbors.h
 */
#include "tf_subsampling/grid_subsampling/grid_subsampling.h"
#include "tf_neighbors/neighbors/neighbors.h"
#include <cstdlib>
#include <cstring>

static std::vector<PointXYZ> to_pts(const float *p, int n)
{
    std::vector<PointXYZ> v((size_t)n);
    for (int i = 0; i < n; i++) v[i] = PointXYZ(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}

extern "C" {

int ref_batch_grid_subsample(const float *points, int n, const int *batches, int b, float dl,
                             float *out_points, int *out_batches)
{
    std::vector<PointXYZ> op = to_pts(points, n), sp;
    std::vector<float> of, sf;
    std::vector<int> oc, sc, ob(batches, batches + b), sb;
    batch_grid_subsampling(op, sp, of, sf, oc, sc, ob, sb, dl);
    for (size_t i = 0; i < sp.size(); i++) { out_points[3 * i] = sp[i].x; out_points[3 * i + 1] = sp[i].y; out_points[3 * i + 2] = sp[i].z; }
    for (int i = 0; i < b; i++) out_batches[i] = sb[i];
    return (int)sp.size();
}

/* single-cloud TF-flavour grid_subsampling with features + one label column */
int ref_grid_subsample_tf(const float *points, int n, const float *features, int fdim, const int *classes,
                          float dl, float *out_points, float *out_features, int *out_classes)
{
    std::vector<PointXYZ> op = to_pts(points, n), sp;
    std::vector<float> of, sf;
    std::vector<int> oc, sc;
    if (features) of.assign(features, features + (size_t)n * fdim);
    if (classes) oc.assign(classes, classes + n);
    grid_subsampling(op, sp, of, sf, oc, sc, dl);
    for (size_t i = 0; i < sp.size(); i++) { out_points[3 * i] = sp[i].x; out_points[3 * i + 1] = sp[i].y; out_points[3 * i + 2] = sp[i].z; }
    if (features) std::memcpy(out_features, sf.data(), sf.size() * sizeof(float));
    if (classes) std::memcpy(out_classes, sc.data(), sc.size() * sizeof(int));
    return (int)sp.size();
}

int *ref_batch_radius_neighbors(const float *queries, int nq, const float *supports, int ns,
                                const int *q_batches, const int *s_batches, int b, float radius,
                                int *max_count_out)
{
    std::vector<PointXYZ> q = to_pts(queries, nq), s = to_pts(supports, ns);
    std::vector<int> qb(q_batches, q_batches + b), sb(s_batches, s_batches + b), out;
    batch_nanoflann_neighbors(q, s, qb, sb, out, radius);
    int mc = nq > 0 ? (int)(out.size() / (size_t)nq) : 0;
    *max_count_out = mc;
    int *r = (int *)std::malloc(sizeof(int) * (out.size() ? out.size() : 1));
    std::memcpy(r, out.data(), out.size() * sizeof(int));
    return r;
}

void ref_free(void *p) { std::free(p); }
}
