// group_ops.cu — a3/a4/a6 stand-alone indexed-gather operators of the pointops API
// (grouping, subtraction, aggregation, interpolation; forward + backward).
// Reference: pytorch/lib/pointops/src/{grouping,subtraction,aggregation,interpolation}/*_cuda_kernel.cu.
// All are HBM/L2-bandwidth bound: rows are moved as float4 when c % 4 == 0, one (row, 4-channel)
// slot per thread so a warp covers contiguous 512 B of a row; per-point reductions that the
// reference does with float atomics (grad wrt the centre point, grad_weight) are plain
// deterministic sums here — only true scatters (grad wrt gathered rows) use atomics.
#include "common.cuh"

static inline int nblocks(long long work, int threads)
{
    long long g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > 2147483647LL) g = 2147483647LL;
    return (int)g;
}

__device__ __forceinline__ void red_add4(float *addr, float4 v)
{
#if __CUDA_ARCH__ >= 900
    atomicAdd(reinterpret_cast<float4 *>(addr), v);
#else
    atomicAdd(addr, v.x); atomicAdd(addr + 1, v.y); atomicAdd(addr + 2, v.z); atomicAdd(addr + 3, v.w);
#endif
}

// ---- grouping (grouping_cuda_kernel.cu:5-25) ----------------------------------------------------
template <int VEC>
__global__ void k_grouping_fwd(long long rows, int cv, const float *__restrict__ input, const int *__restrict__ idx,
                               float *__restrict__ output)
{
    const long long total = rows * cv;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / cv;
        const int col = (int)(e - row * cv);
        const long long src = (long long)__ldg(idx + row) * cv + col;
        if (VEC == 4) reinterpret_cast<float4 *>(output)[e] = __ldg(reinterpret_cast<const float4 *>(input) + src);
        else output[e] = __ldg(input + src);
    }
}

template <int VEC>
__global__ void k_grouping_bwd(long long rows, int cv, const float *__restrict__ grad_output,
                               const int *__restrict__ idx, float *__restrict__ grad_input)
{
    const long long total = rows * cv;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / cv;
        const int col = (int)(e - row * cv);
        const long long dst = (long long)__ldg(idx + row) * cv + col;
        if (VEC == 4) red_add4(grad_input + dst * 4, __ldg(reinterpret_cast<const float4 *>(grad_output) + e));
        else atomicAdd(grad_input + dst, __ldg(grad_output + e));
    }
}

extern "C" int cb_grouping_forward(int m, int nsample, int c, const float *input, const int *idx, float *output, void *stream)
{
    CB_REQUIRE(m >= 0 && nsample >= 0 && c >= 0, CB_EINVAL, "cb_grouping_forward: negative size");
    const long long rows = (long long)m * nsample;
    if (rows == 0 || c == 0) return CB_OK;
    CB_REQUIRE(input && idx && output, CB_EINVAL, "cb_grouping_forward: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (c % 4 == 0) && (((uintptr_t)input | (uintptr_t)output) % 16 == 0);
    if (v4) k_grouping_fwd<4><<<nblocks(rows * (c / 4), 256), 256, 0, st>>>(rows, c / 4, input, idx, output);
    else k_grouping_fwd<1><<<nblocks(rows * c, 256), 256, 0, st>>>(rows, c, input, idx, output);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_grouping_forward");
    return CB_OK;
}

extern "C" int cb_grouping_backward(int m, int nsample, int c, const float *grad_output, const int *idx,
                                    float *grad_input, void *stream)
{
    CB_REQUIRE(m >= 0 && nsample >= 0 && c >= 0, CB_EINVAL, "cb_grouping_backward: negative size");
    const long long rows = (long long)m * nsample;
    if (rows == 0 || c == 0) return CB_OK;
    CB_REQUIRE(grad_output && idx && grad_input, CB_EINVAL, "cb_grouping_backward: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (c % 4 == 0) && (((uintptr_t)grad_input | (uintptr_t)grad_output) % 16 == 0);
    if (v4) k_grouping_bwd<4><<<nblocks(rows * (c / 4), 256), 256, 0, st>>>(rows, c / 4, grad_output, idx, grad_input);
    else k_grouping_bwd<1><<<nblocks(rows * c, 256), 256, 0, st>>>(rows, c, grad_output, idx, grad_input);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_grouping_backward");
    return CB_OK;
}

// ---- subtraction (subtraction_cuda_kernel.cu:5-30) ----------------------------------------------
__global__ void k_subtraction_fwd(long long n, int k, int c, const float *__restrict__ in1, const float *__restrict__ in2,
                                  const int *__restrict__ idx, float *__restrict__ out)
{
    const long long total = n * k * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / c;          // n*k + k_i
        const int ci = (int)(e - row * c);
        const long long ni = row / k;
        out[e] = __ldg(in1 + ni * c + ci) - __ldg(in2 + (long long)__ldg(idx + row) * c + ci);
    }
}

__global__ void k_subtraction_bwd(long long n, int k, int c, const int *__restrict__ idx, const float *__restrict__ go,
                                  float *__restrict__ g1, float *__restrict__ g2)
{
    const long long total = n * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ni = e / c;
        const int ci = (int)(e - ni * c);
        float acc = 0.f;
        for (int ki = 0; ki < k; ki++) {
            const float g = __ldg(go + (ni * k + ki) * c + ci);
            acc += g;
            atomicAdd(g2 + (long long)__ldg(idx + ni * k + ki) * c + ci, -g);
        }
        g1[e] += acc;
    }
}

extern "C" int cb_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2, const int *idx,
                                      float *output, void *stream)
{
    CB_REQUIRE(n >= 0 && nsample >= 0 && c >= 0, CB_EINVAL, "cb_subtraction_forward: negative size");
    if ((long long)n * nsample * c == 0) return CB_OK;
    CB_REQUIRE(input1 && input2 && idx && output, CB_EINVAL, "cb_subtraction_forward: NULL pointer");
    k_subtraction_fwd<<<nblocks((long long)n * nsample * c, 256), 256, 0, (cudaStream_t)stream>>>(n, nsample, c, input1, input2, idx, output);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_subtraction_forward");
    return CB_OK;
}

extern "C" int cb_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output,
                                       float *grad_input1, float *grad_input2, void *stream)
{
    CB_REQUIRE(n >= 0 && nsample >= 0 && c >= 0, CB_EINVAL, "cb_subtraction_backward: negative size");
    if ((long long)n * nsample * c == 0) return CB_OK;
    CB_REQUIRE(idx && grad_output && grad_input1 && grad_input2, CB_EINVAL, "cb_subtraction_backward: NULL pointer");
    k_subtraction_bwd<<<nblocks((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(n, nsample, c, idx, grad_output, grad_input1, grad_input2);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_subtraction_backward");
    return CB_OK;
}

// ---- aggregation (aggregation_cuda_kernel.cu:5-39) ----------------------------------------------
__global__ void k_aggregation_fwd(long long n, int k, int c, int wc, const float *__restrict__ input,
                                  const float *__restrict__ pos, const float *__restrict__ w, const int *__restrict__ idx,
                                  float *__restrict__ out)
{
    const long long total = n * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ni = e / c;
        const int ci = (int)(e - ni * c);
        const int wi = ci % wc;
        float acc = 0.f;
        for (int ki = 0; ki < k; ki++) {
            const long long r = ni * k + ki;
            const float a = __ldg(input + (long long)__ldg(idx + r) * c + ci) + __ldg(pos + r * c + ci);
            acc = __fmaf_rn(a, __ldg(w + r * wc + wi), acc);      // same contraction as the reference's SASS
        }
        out[e] = acc;
    }
}

// grad_input scatter + grad_position
__global__ void k_aggregation_bwd_ip(long long n, int k, int c, int wc, const float *__restrict__ w,
                                     const int *__restrict__ idx, const float *__restrict__ go, float *__restrict__ gi,
                                     float *__restrict__ gp)
{
    const long long total = n * k * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / c;
        const int ci = (int)(e - r * c);
        const long long ni = r / k;
        const float g = __ldg(go + ni * c + ci) * __ldg(w + r * wc + ci % wc);
        gp[e] = g;
        atomicAdd(gi + (long long)__ldg(idx + r) * c + ci, g);
    }
}

// grad_weight[n,k,j] = sum_{ci % wc == j} go[n,ci] * (input[idx,ci] + pos[n,k,ci])   (no atomics)
__global__ void k_aggregation_bwd_w(long long n, int k, int c, int wc, const float *__restrict__ input,
                                    const float *__restrict__ pos, const int *__restrict__ idx,
                                    const float *__restrict__ go, float *__restrict__ gw)
{
    const long long total = n * k * wc;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / wc;
        const int j = (int)(e - r * wc);
        const long long ni = r / k;
        const long long src = (long long)__ldg(idx + r) * c;
        float acc = 0.f;
        for (int ci = j; ci < c; ci += wc)
            acc += __ldg(go + ni * c + ci) * (__ldg(input + src + ci) + __ldg(pos + r * c + ci));
        gw[e] += acc;
    }
}

extern "C" int cb_aggregation_forward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                      const float *weight, const int *idx, float *output, void *stream)
{
    CB_REQUIRE(n >= 0 && nsample >= 0 && c >= 0 && w_c > 0, CB_EINVAL, "cb_aggregation_forward: bad size");
    if ((long long)n * c == 0) return CB_OK;
    CB_REQUIRE(input && position && weight && idx && output, CB_EINVAL, "cb_aggregation_forward: NULL pointer");
    k_aggregation_fwd<<<nblocks((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(n, nsample, c, w_c, input, position, weight, idx, output);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_aggregation_forward");
    return CB_OK;
}

extern "C" int cb_aggregation_backward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                       const float *weight, const int *idx, const float *grad_output, float *grad_input,
                                       float *grad_position, float *grad_weight, void *stream)
{
    CB_REQUIRE(n >= 0 && nsample >= 0 && c >= 0 && w_c > 0, CB_EINVAL, "cb_aggregation_backward: bad size");
    if ((long long)n * nsample * c == 0) return CB_OK;
    CB_REQUIRE(input && position && weight && idx && grad_output && grad_input && grad_position && grad_weight, CB_EINVAL,
               "cb_aggregation_backward: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    k_aggregation_bwd_ip<<<nblocks((long long)n * nsample * c, 256), 256, 0, st>>>(n, nsample, c, w_c, weight, idx, grad_output, grad_input, grad_position);
    k_aggregation_bwd_w<<<nblocks((long long)n * nsample * w_c, 256), 256, 0, st>>>(n, nsample, c, w_c, input, position, idx, grad_output, grad_weight);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_aggregation_backward");
    return CB_OK;
}

// ---- interpolation (interpolation_cuda_kernel.cu:5-33) -------------------------------------------
__global__ void k_interpolation_fwd(long long n, int c, int k, const float *__restrict__ input, const int *__restrict__ idx,
                                    const float *__restrict__ w, float *__restrict__ out)
{
    const long long total = n * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ni = e / c;
        const int ci = (int)(e - ni * c);
        float acc = 0.f;
        for (int i = 0; i < k; i++)
            acc = __fmaf_rn(__ldg(input + (long long)__ldg(idx + ni * k + i) * c + ci), __ldg(w + ni * k + i), acc);
        out[e] = acc;
    }
}

__global__ void k_interpolation_bwd(long long n, int c, int k, const float *__restrict__ go, const int *__restrict__ idx,
                                    const float *__restrict__ w, float *__restrict__ gi)
{
    const long long total = n * c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ni = e / c;
        const int ci = (int)(e - ni * c);
        const float g = __ldg(go + e);
        for (int i = 0; i < k; i++)
            atomicAdd(gi + (long long)__ldg(idx + ni * k + i) * c + ci, g * __ldg(w + ni * k + i));
    }
}

extern "C" int cb_interpolation_forward(int n, int c, int k, const float *input, const int *idx, const float *weight,
                                        float *output, void *stream)
{
    CB_REQUIRE(n >= 0 && c >= 0 && k >= 0, CB_EINVAL, "cb_interpolation_forward: negative size");
    if ((long long)n * c == 0) return CB_OK;
    CB_REQUIRE(input && idx && weight && output, CB_EINVAL, "cb_interpolation_forward: NULL pointer");
    k_interpolation_fwd<<<nblocks((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(n, c, k, input, idx, weight, output);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_interpolation_forward");
    return CB_OK;
}

extern "C" int cb_interpolation_backward(int n, int c, int k, const float *grad_output, const int *idx, const float *weight,
                                         float *grad_input, void *stream)
{
    CB_REQUIRE(n >= 0 && c >= 0 && k >= 0, CB_EINVAL, "cb_interpolation_backward: negative size");
    if ((long long)n * c == 0) return CB_OK;
    CB_REQUIRE(grad_output && idx && weight && grad_input, CB_EINVAL, "cb_interpolation_backward: NULL pointer");
    k_interpolation_bwd<<<nblocks((long long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(n, c, k, grad_output, idx, weight, grad_input);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_interpolation_backward");
    return CB_OK;
}
