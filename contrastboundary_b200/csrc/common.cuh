// common.cuh — shared helpers for libcbops (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cbops.h"

#define CB_FULL_MASK 0xffffffffu

void cb_set_error(const char *fmt, ...);
extern unsigned long long g_cb_launches;   // kernels launched by this library (host-side count)
#define CB_COUNT(n) ((void)__atomic_fetch_add(&g_cb_launches, (unsigned long long)(n), __ATOMIC_RELAXED))   // autograd / DDP threads launch too

#define CB_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            cb_set_error(__VA_ARGS__);         \
            return (code);                     \
        }                                      \
    } while (0)

#define CB_CUDA_CHECK(what)                                                             \
    do {                                                                                \
        cudaError_t e__ = cudaPeekAtLastError();                                        \
        if (e__ != cudaSuccess) {                                                       \
            cb_set_error("%s: %s", what, cudaGetErrorString(e__));                      \
            (void)cudaGetLastError();                                                   \
            return CB_ECUDA;                                                            \
        }                                                                               \
    } while (0)

static inline size_t cb_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A step of this network is a chain of ~1 300 short kernels; between two dependent kernels
// of a stream (or of a captured graph) the GPU idles ~2.5 us while the next grid is set up.  A kernel launched through
// cb_launch_pdl may be SET UP while its predecessor still runs; its first statement, cb_pdl_wait(), blocks until every
// grid it depends on has completed and flushed its memory, so the semantics are those of an ordinary in-order launch.
// Only kernels that start with cb_pdl_wait() may be launched this way.  cb_set_pdl(0) / CB_PDL=0 = ordinary launches.
// ---------------------------------------------------------------------------------------------
extern int g_cb_pdl;
__device__ __forceinline__ void cb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cb_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t cb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_cb_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Squared distance exactly as the reference's SASS computes it (nvcc -fmad=true on
// (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f)):  t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t)
// (knnquery_cuda_kernel.cu:99, sampling_cuda_kernel.cu:54; order read off the reference's sm_100a SASS
// and pinned by tests/golden/pointops_ref_gpu.npz).
__device__ __forceinline__ float cb_sqdist(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    t = __fmaf_rn(dz, dz, t);
    return t;
}

// first i with q < offset[i]  (knnquery_cuda_kernel.cu:51-62), by binary search
__device__ __forceinline__ int cb_scene_of(int q, const int *__restrict__ offset, int b)
{
    int lo = 0, hi = b - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (q < __ldg(offset + mid)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ unsigned cb_f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float cb_ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
