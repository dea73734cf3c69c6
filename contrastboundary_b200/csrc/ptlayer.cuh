// ptlayer.cuh — shared device helpers for the fused PointTransformer local aggregation.
#pragma once
#include "common.cuh"

// Channel mapping of the "c-space" kernels: a warp owns whole feature rows; lane `l` owns
//   C >= 128 : channels s*128 + 4*l + v   (s < NS = C/128 slices, v < 4)  -> one LDG.128 per slice
//   C == 64  : channels 2*l + v           (v < 2)
//   C == 32  : channel  l
template <int C>
struct PtMap {
    static constexpr int VW = C >= 128 ? 4 : C / 32;
    static constexpr int NS = C >= 128 ? C / 128 : 1;
    static constexpr int CS = C / 8;                 // share_planes = 8 (blocks.py:14)
    static constexpr int L = CS / VW < 32 ? CS / VW : 32;   // lanes with equal (lane % L) share j = ch % CS
    __device__ static __forceinline__ int ch(int lane, int s, int v) { return s * 128 + lane * VW + v; }
};

template <int VW>
__device__ __forceinline__ void pt_load(const float *__restrict__ p, float (&v)[VW])
{
    if (VW == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if (VW == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VW>
__device__ __forceinline__ void pt_store(float *p, const float (&v)[VW])
{
    if (VW == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if (VW == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    else p[0] = v[0];
}

template <int VW>
__device__ __forceinline__ void pt_red_add(float *p, const float (&v)[VW])
{
    if (VW == 4) {
        atomicAdd(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else if (VW == 2) {
        atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
    } else {
        atomicAdd(p, v[0]);
    }
}

// batch-norm affine of one channel: y = x * sc + sh  (sc = gamma * invstd, sh = beta - mean * sc)
struct PtBn {
    const float *scale, *shift, *mean, *invstd;
};
