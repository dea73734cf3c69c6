// fps.cu — a2: farthest point sampling, bit-identical to the reference
// (pytorch/lib/pointops/src/sampling/sampling_cuda_kernel.cu:14-171).
//
// The reference runs ONE block per scene and re-reads xyz and the running min-distance buffer
// from global memory in each of its m-1 strictly sequential iterations.  Here each scene is owned
// by a thread-block CLUSTER: every thread keeps its points and their running min-distance in
// registers for the whole kernel; one iteration = per-thread update -> warp arg-max (REDUX) ->
// CTA arg-max through shared memory -> cluster arg-max through distributed shared memory.
//
// Tie rule.  The reference's winner among equal maxima is fixed by its launch geometry: thread
// t = r % B of a B-thread block (B = opt_n_threads(n_max), cuda_utils.h:11-14) scans r = t, t+B, ...
// keeping the first strict maximum, and the shared-memory tree keeps the lower slot unless the upper
// is strictly greater (sampling_cuda_kernel.cu:5-10) — so the winner has the smallest
// (bit-reverse_B(t), r / B).  We give every point that 32-bit priority and reduce
// the pair (distance bits, ~priority) with max, which reproduces the rule under any reduction order.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
namespace cg = cooperative_groups;

#define FPS_THREADS 1024

struct FpsCand {
    unsigned d;   // distance bits (non-negative floats order like unsigned ints)
    unsigned np;  // ~priority (larger wins)
};

__device__ __forceinline__ bool fps_better(unsigned d1, unsigned p1, unsigned d0, unsigned p0)
{
    return d1 > d0 || (d1 == d0 && p1 > p0);
}

template <int CL, int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1) k_fps(const float *__restrict__ xyz, const int *__restrict__ offset,
                                                        const int *__restrict__ new_offset, float *__restrict__ tmp,
                                                        int *__restrict__ idx, int logB)
{
    __shared__ FpsCand s_warp[2][32];
    __shared__ FpsCand s_xchg[2][CL > 1 ? CL : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const int bid = blockIdx.x / CL;
    const int start_n = bid == 0 ? 0 : offset[bid - 1], end_n = offset[bid];
    const int start_m = bid == 0 ? 0 : new_offset[bid - 1], end_m = new_offset[bid];
    const int ns = end_n - start_n;
    const int B = 1 << logB;
    const unsigned S = (unsigned)((ns + B - 1) >> logB);     // scan steps per reference thread

    constexpr int NP = PPT > 0 ? PPT : 1;
    float px[NP], py[NP], pz[NP], md[NP];
    unsigned npri[NP];
    const int stride = CL * FPS_THREADS;
    const int r0 = rank * FPS_THREADS + tid;
    if (PPT > 0) {
#pragma unroll
        for (int u = 0; u < NP; u++) {
            const int r = r0 + u * stride;
            if (r < ns) {
                const int k = start_n + r;
                px[u] = xyz[3 * k]; py[u] = xyz[3 * k + 1]; pz[u] = xyz[3 * k + 2];
                md[u] = tmp[k];
                const unsigned t = (unsigned)r & (unsigned)(B - 1);
                const unsigned brev = logB ? (__brev(t) >> (32 - logB)) : 0u;
                npri[u] = ~(brev * S + ((unsigned)r >> logB));
            } else {
                px[u] = py[u] = pz[u] = 0.f; md[u] = -1.f; npri[u] = 0u;
            }
        }
    }
    if (ns <= 0 || end_m <= start_m) return;   // uniform per cluster
    if (rank == 0 && tid == 0) idx[start_m] = start_n;       // sampling_cuda_kernel.cu:39
    int old = start_n;
    for (int j = start_m + 1; j < end_m; j++) {
        const int par = j & 1;
        const float ox = __ldg(xyz + 3 * old), oy = __ldg(xyz + 3 * old + 1), oz = __ldg(xyz + 3 * old + 2);
        unsigned bd = 0u, bp = 0u;    // (0, 0) loses to every real point (np of real points is >= ~(B*S) > 0)
        if (PPT > 0) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                const float d = cb_sqdist(px[u], py[u], pz[u], ox, oy, oz);
                const float m2 = fminf(d, md[u]);
                md[u] = m2;
                const unsigned db = __float_as_uint(fmaxf(m2, 0.f));
                const bool real = npri[u] != 0u;
                if (real && fps_better(db, npri[u], bd, bp)) { bd = db; bp = npri[u]; }
            }
        } else {
            for (int r = r0; r < ns; r += stride) {
                const int k = start_n + r;
                const float d = cb_sqdist(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], ox, oy, oz);
                const float m2 = fminf(d, tmp[k]);
                tmp[k] = m2;
                const unsigned t = (unsigned)r & (unsigned)(B - 1);
                const unsigned brev = logB ? (__brev(t) >> (32 - logB)) : 0u;
                const unsigned np = ~(brev * S + ((unsigned)r >> logB));
                const unsigned db = __float_as_uint(fmaxf(m2, 0.f));
                if (fps_better(db, np, bd, bp)) { bd = db; bp = np; }
            }
        }
        // warp arg-max
        unsigned wd = __reduce_max_sync(CB_FULL_MASK, bd);
        unsigned wp = __reduce_max_sync(CB_FULL_MASK, bd == wd ? bp : 0u);
        if (lane == 0) { s_warp[par][warp].d = wd; s_warp[par][warp].np = wp; }
        __syncthreads();
        // CTA arg-max (every warp redundantly)
        const FpsCand c = s_warp[par][lane];
        unsigned cd = __reduce_max_sync(CB_FULL_MASK, c.d);
        unsigned cp = __reduce_max_sync(CB_FULL_MASK, c.d == cd ? c.np : 0u);
        if (CL > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (warp == 0 && lane < CL) {
                FpsCand *dst = cluster.map_shared_rank(&s_xchg[par][rank], lane);
                dst->d = cd; dst->np = cp;
            }
            cluster.sync();
            const FpsCand x = s_xchg[par][lane < CL ? lane : 0];
            cd = __reduce_max_sync(CB_FULL_MASK, x.d);
            cp = __reduce_max_sync(CB_FULL_MASK, x.d == cd ? x.np : 0u);
        }
        // priority -> point index
        const unsigned pri = ~cp;
        const unsigned brev = pri / S, step = pri - brev * S;
        const unsigned t = logB ? (__brev(brev) >> (32 - logB)) : 0u;
        old = start_n + (int)((step << logB) + t);
        if (rank == 0 && tid == 0) idx[j] = old;
    }
    if (PPT > 0) {   // leave the running min-distance where the reference leaves it
#pragma unroll
        for (int u = 0; u < NP; u++) {
            const int r = r0 + u * stride;
            if (r < ns) tmp[start_n + r] = md[u];
        }
    }
}

template <int CL, int PPT>
static cudaError_t launch_fps(int b, const float *xyz, const int *offset, const int *new_offset, float *tmp, int *idx,
                              int logB, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * CL));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_fps<CL, PPT>, xyz, offset, new_offset, tmp, idx, logB);
}

extern "C" int cb_furthest_sampling(int b, int n_max, const float *xyz, const int *offset, const int *new_offset,
                                    float *tmp, int *idx, void *stream)
{
    CB_REQUIRE(b >= 0 && n_max >= 0, CB_EINVAL, "cb_furthest_sampling: negative size");
    if (b == 0 || n_max == 0) return CB_OK;
    CB_REQUIRE(xyz && offset && new_offset && tmp && idx, CB_EINVAL, "cb_furthest_sampling: NULL pointer");
    // reference block size: largest power of two <= n_max, capped at 1024 (cuda_utils.h:11-14)
    // (the same double-precision expression, so that exact powers of two round the same way)
    int pow_2 = (int)(log((double)n_max) / log(2.0));
    if (pow_2 > 10) pow_2 = 10;
    if (pow_2 < 0) pow_2 = 0;
    const int logB = pow_2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (n_max <= 1024) e = launch_fps<1, 1>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 2048) e = launch_fps<1, 2>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 4096) e = launch_fps<1, 4>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 8192) e = launch_fps<1, 8>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 16384) e = launch_fps<8, 2>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 40960) e = launch_fps<8, 5>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else if (n_max <= 65536) e = launch_fps<8, 8>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    else e = launch_fps<8, 0>(b, xyz, offset, new_offset, tmp, idx, logB, st);
    if (e != cudaSuccess) {
        cb_set_error("cb_furthest_sampling: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return CB_ECUDA;
    }
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_furthest_sampling");
    return CB_OK;
}
