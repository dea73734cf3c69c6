// tf32.cuh — 3xTF32 helpers for the tensor-core kernels (tc_gemm.cu, ptlayer_mma.cu).
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi); a*b ~= hi*hi + lo*hi + hi*lo with FP32 accumulation keeps the
// product error near 2^-21 relative (the lo*lo term is below FP32 rounding).
#pragma once

__device__ __forceinline__ unsigned tg_tf32(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tg_split(float x, unsigned &hi, unsigned &lo)
{
    hi = tg_tf32(x);
    lo = tg_tf32(x - __uint_as_float(hi));
}
// D += A(16x8, row) * B(8x8, col);  fragment layout (g = lane >> 2, t = lane & 3):
//   a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4);  b0 (k t, n g)  b1 (k t+4, n g)
//   d0 (row g, n 2t)  d1 (row g, n 2t+1)  d2 (row g+8, n 2t)  d3 (row g+8, n 2t+1)
__device__ __forceinline__ void tg_mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
