// bn_ops.cu — fused BatchNorm1d (+ residual) (+ ReLU) over per-point feature matrices (n, C), forward and backward.
// Reference: the `relu(bn(linear(x)))` / `relu(bn3(linear3(x)) + identity)` patterns of PointTransformerBlock,
// TransitionDown/Up and the head MLPs (pytorch/model/blocks.py:76,104-108,125-133,157-192), which the reference runs
// as separate BatchNorm, add and ReLU kernels (forward: statistics, running-stat update, normalise, add, relu; backward:
// relu, reduce, elementwise).  Here: forward = column statistics + ONE apply kernel, backward = ONE reduce + ONE apply.
//   forward : y = act( (x - mean) * invstd * gamma + beta [+ residual] )     act = relu | identity
//   backward: dy = gy * [y > 0];  dx = gamma*invstd * (dy - mean(dy) - xhat * mean(dy*xhat));  dres = dy;
//             dgamma = sum dy*xhat;  dbeta = sum dy
// Training mode uses the batch statistics (biased variance) and updates the running statistics with the unbiased
// variance, exactly as torch.nn.BatchNorm1d; evaluation mode uses the running statistics.
// Statistics are accumulated in double (one atomic per column and block), so the result does not depend on n.
#include "common.cuh"

#define BN_THREADS 256

// per-column sums of a (sum x, sum x^2) or of (dy, dy * xhat) — rows strided over blocks and row-lanes
// MODE 0: forward statistics of x.   MODE 1: backward sums; dy = relu ? gy * [y > 0] : gy, xhat = (x - mean) * invstd
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) k_bn_colsums(long long n, int C, const float *__restrict__ x,
                                                           const float *__restrict__ gy, const float *__restrict__ y,
                                                           const float *__restrict__ bnbuf, int relu,
                                                           double *__restrict__ sums)
{
    extern __shared__ float bs_sm[];            // [2][C]
    cb_pdl_wait();
    for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS) bs_sm[i] = 0.f;
    __syncthreads();
    // thread -> 4 consecutive columns (C % 4 == 0), row lanes = BN_THREADS / (C / 4) (>= 1 for C <= 1024)
    const int cq = C >> 2;
    const int lanes = BN_THREADS / cq > 0 ? BN_THREADS / cq : 1;
    const int colq = threadIdx.x % cq, rl = threadIdx.x / cq;
    if (rl < lanes) {
        float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
        float mean[4] = {0, 0, 0, 0}, inv[4] = {0, 0, 0, 0};
        if (MODE == 1) {
#pragma unroll
            for (int e = 0; e < 4; e++) { mean[e] = bnbuf[2 * C + 4 * colq + e]; inv[e] = bnbuf[3 * C + 4 * colq + e]; }
        }
        for (long long r = (long long)blockIdx.x * lanes + rl; r < n; r += (long long)gridDim.x * lanes) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + r * C) + colq);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
            if (MODE == 0) {
#pragma unroll
                for (int e = 0; e < 4; e++) { s1[e] += xa[e]; s2[e] += xa[e] * xa[e]; }
            } else {
                const float4 gv = __ldg(reinterpret_cast<const float4 *>(gy + r * C) + colq);
                float ga[4] = {gv.x, gv.y, gv.z, gv.w};
                if (relu) {
                    const float4 yv = __ldg(reinterpret_cast<const float4 *>(y + r * C) + colq);
                    const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) ga[e] = ya[e] > 0.f ? ga[e] : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 4; e++) { s1[e] += ga[e]; s2[e] += ga[e] * ((xa[e] - mean[e]) * inv[e]); }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) { atomicAdd(&bs_sm[4 * colq + e], s1[e]); atomicAdd(&bs_sm[C + 4 * colq + e], s2[e]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS) atomicAdd(sums + i, (double)bs_sm[i]);
}

// forward apply: every block derives scale / shift from the statistics; block 0 also writes bnbuf and the running stats
// Self-cleaning accumulators (ticket != NULL): the statistics buffer is PERSISTENT and zero between uses — the column-sum
// kernel accumulates into it, every block of the apply kernel reads it in its prologue and then takes a ticket; the block
// that takes the last ticket knows everyone has read, and zeroes buffer and ticket for the next use.  This removes the
// cudaMemsetAsync in front of every BatchNorm (2 per layer and step: 138 graph nodes of this network).
__device__ __forceinline__ void bn_self_clean(double *acc, int count, int *ticket)
{
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        for (int c = threadIdx.x; c < count; c += blockDim.x) acc[c] = 0.0;
        if (threadIdx.x == 0) *ticket = 0;
    }
}

__global__ void __launch_bounds__(BN_THREADS) k_bn_apply(long long n, int C, const float *__restrict__ x,
                                                         const float *__restrict__ residual, double *stats,
                                                         const float *__restrict__ gamma, const float *__restrict__ beta,
                                                         float *running_mean, float *running_var, float momentum, float eps,
                                                         int training, int relu, float *__restrict__ y, float *__restrict__ bnbuf,
                                                         int *ticket)
{
    extern __shared__ float ba_sm[];            // [2][C]: scale, shift
    cb_pdl_wait();
    const double count = (double)n;
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        float mean, var;
        double unbiased = 0.0;
        if (training) {
            const double m = stats[c] / count;
            double v = stats[C + c] / count - m * m;
            if (v < 0) v = 0;
            mean = (float)m; var = (float)v;
            unbiased = count > 1 ? v * count / (count - 1) : v;
        } else {
            mean = running_mean[c]; var = running_var[c];
        }
        const float invstd = 1.0f / sqrtf(var + eps);
        const float sc = gamma[c] * invstd;
        ba_sm[c] = sc; ba_sm[C + c] = beta[c] - mean * sc;
        if (blockIdx.x == 0) {
            bnbuf[c] = sc; bnbuf[C + c] = beta[c] - mean * sc; bnbuf[2 * C + c] = mean; bnbuf[3 * C + c] = invstd;
            if (training) {
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        }
    }
    __syncthreads();
    if (ticket && training) bn_self_clean(stats, 2 * C, ticket);
    const int cq = C >> 2;
    const long long total = n * cq;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BN_THREADS) {
        const int colq = (int)(i % cq);
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + i);
        const float4 sc = *reinterpret_cast<const float4 *>(ba_sm + 4 * colq), sh = *reinterpret_cast<const float4 *>(ba_sm + C + 4 * colq);
        float4 o = make_float4(xv.x * sc.x + sh.x, xv.y * sc.y + sh.y, xv.z * sc.z + sh.z, xv.w * sc.w + sh.w);
        if (residual) {
            const float4 rv = __ldg(reinterpret_cast<const float4 *>(residual) + i);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        reinterpret_cast<float4 *>(y)[i] = o;
    }
}

// backward apply: dx = kk * (dy - ma - xhat * mb); dres = dy; block 0 writes dgamma / dbeta
__global__ void __launch_bounds__(BN_THREADS) k_bn_bwd_apply(long long n, int C, const float *__restrict__ x,
                                                             const float *__restrict__ gy, const float *__restrict__ y,
                                                             const float *__restrict__ gamma, const float *__restrict__ bnbuf,
                                                             double *sums, int training, int relu,
                                                             float *__restrict__ gx, float *__restrict__ gres,
                                                             float *__restrict__ dgamma, float *__restrict__ dbeta, int *ticket)
{
    extern __shared__ float bb_sm[];            // [5][C]: kk, ma, mb, mean, invstd
    cb_pdl_wait();
    const double count = (double)n;
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        const double sa = sums[c], sb = sums[C + c];
        const float invstd = bnbuf[3 * C + c];
        bb_sm[c] = gamma[c] * invstd;
        bb_sm[C + c] = training ? (float)(sa / count) : 0.f;
        bb_sm[2 * C + c] = training ? (float)(sb / count) : 0.f;
        bb_sm[3 * C + c] = bnbuf[2 * C + c];
        bb_sm[4 * C + c] = invstd;
        if (blockIdx.x == 0) { dgamma[c] = (float)sb; dbeta[c] = (float)sa; }
    }
    __syncthreads();
    if (ticket) bn_self_clean(sums, 2 * C, ticket);
    const int cq = C >> 2;
    const long long total = n * cq;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BN_THREADS) {
        const int c0 = 4 * (int)(i % cq);
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + i);
        const float4 gv = __ldg(reinterpret_cast<const float4 *>(gy) + i);
        float ga[4] = {gv.x, gv.y, gv.z, gv.w};
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
        if (relu) {
            const float4 yv = __ldg(reinterpret_cast<const float4 *>(y) + i);
            const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int e = 0; e < 4; e++) ga[e] = ya[e] > 0.f ? ga[e] : 0.f;
        }
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int c = c0 + e;
            const float xhat = (xa[e] - bb_sm[3 * C + c]) * bb_sm[4 * C + c];
            o[e] = bb_sm[c] * (ga[e] - bb_sm[C + c] - xhat * bb_sm[2 * C + c]);
        }
        reinterpret_cast<float4 *>(gx)[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (gres) reinterpret_cast<float4 *>(gres)[i] = make_float4(ga[0], ga[1], ga[2], ga[3]);
    }
}

static int bn_grid(long long work_items, int per_block)
{
    long long g = (work_items + per_block - 1) / per_block;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" int cb_bn_act_forward(long long n, int c, const float *x, const float *residual, const float *gamma,
                                 const float *beta, float *running_mean, float *running_var, float momentum, float eps,
                                 int training, int relu, float *y, float *bnbuf, double *stats, void *stream)
{
    CB_REQUIRE(n >= 0 && c > 0 && c % 4 == 0 && c <= 1024, CB_EINVAL, "cb_bn_act_forward: n=%lld c=%d (c %% 4 == 0, c <= 1024)", n, c);
    CB_REQUIRE(x && gamma && beta && running_mean && running_var && y && bnbuf && stats, CB_EINVAL, "cb_bn_act_forward: NULL pointer");
    CB_REQUIRE(((((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual)) & 15) == 0, CB_EINVAL, "cb_bn_act_forward: 16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return CB_OK;
    const bool persistent = (training & 2) != 0;       // stats: 2c doubles + an int ticket, zero on entry, zero again on exit
    training &= 1;
    int *ticket = persistent ? reinterpret_cast<int *>(stats + 2 * c) : nullptr;
    if (training) {
        if (!persistent) cudaMemsetAsync(stats, 0, sizeof(double) * 2 * c, st);
        const int lanes = BN_THREADS / (c / 4) > 0 ? BN_THREADS / (c / 4) : 1;
        cb_launch_pdl(k_bn_colsums<0>, dim3(bn_grid(n, lanes * 8)), dim3(BN_THREADS), 2 * c * sizeof(float), st, n, c, x, nullptr, nullptr,
                      nullptr, 0, stats);
    }
    cb_launch_pdl(k_bn_apply, dim3(bn_grid(n * (c / 4), BN_THREADS * 4)), dim3(BN_THREADS), 2 * c * sizeof(float), st, n, c, x, residual,
                  stats, gamma, beta, running_mean, running_var, momentum, eps, training, relu, y, bnbuf, ticket);
    CB_COUNT(training ? 2 : 1);
    CB_CUDA_CHECK("cb_bn_act_forward");
    return CB_OK;
}

extern "C" int cb_bn_act_backward(long long n, int c, const float *x, const float *y, const float *gamma, const float *bnbuf,
                                  int training, int relu, const float *grad_y, float *grad_x, float *grad_residual,
                                  float *grad_gamma, float *grad_beta, double *sums, void *stream)
{
    CB_REQUIRE(n >= 0 && c > 0 && c % 4 == 0 && c <= 1024, CB_EINVAL, "cb_bn_act_backward: n=%lld c=%d", n, c);
    CB_REQUIRE(x && gamma && bnbuf && grad_y && grad_x && grad_gamma && grad_beta && sums && (y || !relu), CB_EINVAL,
               "cb_bn_act_backward: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool persistent = (training & 2) != 0;       // sums: 2c doubles + an int ticket, zero on entry, zero again on exit
    training &= 1;
    int *ticket = persistent ? reinterpret_cast<int *>(sums + 2 * c) : nullptr;
    if (!persistent) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * c, st);
    if (n == 0) {
        cudaMemsetAsync(grad_gamma, 0, sizeof(float) * c, st);
        cudaMemsetAsync(grad_beta, 0, sizeof(float) * c, st);
        return CB_OK;
    }
    const int lanes = BN_THREADS / (c / 4) > 0 ? BN_THREADS / (c / 4) : 1;
    cb_launch_pdl(k_bn_colsums<1>, dim3(bn_grid(n, lanes * 8)), dim3(BN_THREADS), 2 * c * sizeof(float), st, n, c, x, grad_y, y, bnbuf, relu,
                  sums);
    cb_launch_pdl(k_bn_bwd_apply, dim3(bn_grid(n * (c / 4), BN_THREADS * 4)), dim3(BN_THREADS), 5 * c * sizeof(float), st, n, c, x, grad_y, y,
                  gamma, bnbuf, sums, training, relu, grad_x, grad_residual, grad_gamma, grad_beta, ticket);
    CB_COUNT(2);
    CB_CUDA_CHECK("cb_bn_act_backward");
    return CB_OK;
}
