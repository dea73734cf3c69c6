/*
 * oracle/ref_shim/ref_cpy_shim.cpp — TEST INFRASTRUCTURE.
 * extern "C" entry point over the REFERENCE's CPython-flavour grid_subsampling core
 * (cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp, features mean + per-column
 * label vote), compiled unmodified into oracle/_ref/libref_cpy.so.  Separate .so because it has
 * the same function name as the TF flavour with an extra `int verbose` argument.
 */
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"
#include <cstring>

extern "C" int ref_grid_subsample(const float *points, int n, const float *features, int fdim,
                                  const int *classes, int ldim, float dl, float *out_points,
                                  float *out_features, int *out_classes)
{
    std::vector<PointXYZ> op((size_t)n), sp;
    for (int i = 0; i < n; i++) op[i] = PointXYZ(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    std::vector<float> of, sf;
    std::vector<int> oc, sc;
    if (features) of.assign(features, features + (size_t)n * fdim);
    if (classes) oc.assign(classes, classes + (size_t)n * ldim);
    grid_subsampling(op, sp, of, sf, oc, sc, dl, 0);
    for (size_t i = 0; i < sp.size(); i++) { out_points[3 * i] = sp[i].x; out_points[3 * i + 1] = sp[i].y; out_points[3 * i + 2] = sp[i].z; }
    if (features) std::memcpy(out_features, sf.data(), sf.size() * sizeof(float));
    if (classes) std::memcpy(out_classes, sc.data(), sc.size() * sizeof(int));
    return (int)sp.size();
}
