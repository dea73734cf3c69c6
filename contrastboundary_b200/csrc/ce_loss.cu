// ce_loss.cu — the segmentation cross-entropy of the reference's Loss (nn.CrossEntropyLoss(ignore_index), mean over the
// points that are not ignored; pytorch/model/pointtransformer_seg.py:15-25, tool/train.py:149) in one kernel per direction.
// torch runs it as log_softmax + nll_loss: the nll reductions alone cost 140 + 94 us per step for 163 840 x 13 logits
// (single-block reduce kernels); here a thread owns a row (c <= 1024 classes, 13 in S3DIS), the block reduces in shared
// memory and adds two doubles; the last block (ticket) divides.
#include "common.cuh"

#define CE_THREADS 256

// acc: [0] sum of row losses, [1] number of counted rows, [2] ticket (as int) — zero on entry
__global__ void __launch_bounds__(CE_THREADS) k_ce_forward(int n, int C, const float *__restrict__ logits,
                                                          const long long *__restrict__ target, long long ignore_index,
                                                          double *acc, float *__restrict__ loss)
{
    __shared__ double s_sum[CE_THREADS / 32], s_cnt[CE_THREADS / 32];
    double my = 0.0, cnt = 0.0;
    for (int r = blockIdx.x * CE_THREADS + threadIdx.x; r < n; r += gridDim.x * CE_THREADS) {
        const long long t = target[r];
        if (t == ignore_index || t < 0 || t >= C) continue;
        const float *x = logits + (size_t)r * C;
        float m = x[0];
        for (int c = 1; c < C; c++) m = fmaxf(m, x[c]);
        float se = 0.f;
        for (int c = 0; c < C; c++) se += expf(x[c] - m);
        my += (double)(m + logf(se) - x[t]);
        cnt += 1.0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        my += __shfl_down_sync(0xffffffffu, my, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = my; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < CE_THREADS / 32; w++) { a += s_sum[w]; b += s_cnt[w]; }
        atomicAdd(acc, a);
        atomicAdd(acc + 1, b);
        __threadfence();
        int *ticket = reinterpret_cast<int *>(acc + 2);
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            const double s = atomicAdd(acc, 0.0), k = atomicAdd(acc + 1, 0.0);     // every block's contribution is in
            *loss = (float)(s / k);                                                // 0 / 0 = NaN, as torch
        }
    }
}

// dlogits[r, :] = (softmax(logits[r, :]) - onehot(target[r])) * gout / count, 0 for ignored rows
__global__ void __launch_bounds__(CE_THREADS) k_ce_backward(int n, int C, const float *__restrict__ logits,
                                                           const long long *__restrict__ target, long long ignore_index,
                                                           const double *__restrict__ acc, const float *__restrict__ gout,
                                                           float *__restrict__ dlogits)
{
    const float scale = gout[0] / (float)acc[1];
    for (int r = blockIdx.x * CE_THREADS + threadIdx.x; r < n; r += gridDim.x * CE_THREADS) {
        const long long t = target[r];
        const float *x = logits + (size_t)r * C;
        float *d = dlogits + (size_t)r * C;
        if (t == ignore_index || t < 0 || t >= C) {
            for (int c = 0; c < C; c++) d[c] = 0.f;
            continue;
        }
        float m = x[0];
        for (int c = 1; c < C; c++) m = fmaxf(m, x[c]);
        float se = 0.f;
        for (int c = 0; c < C; c++) se += expf(x[c] - m);
        const float inv = 1.f / se;
        for (int c = 0; c < C; c++) d[c] = (expf(x[c] - m) * inv - (c == (int)t ? 1.f : 0.f)) * scale;
    }
}

static int ce_grid(int n)
{
    int g = (n + CE_THREADS - 1) / CE_THREADS;
    if (g > 148 * 4) g = 148 * 4;
    return g < 1 ? 1 : g;
}

extern "C" int cb_cross_entropy_forward(int n, int c, const float *logits, const long long *target, long long ignore_index,
                                        double *acc, float *loss, void *stream)
{
    CB_REQUIRE(n >= 0 && c > 0 && c <= 1024 && logits && target && acc && loss, CB_EINVAL, "cb_cross_entropy_forward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
    k_ce_forward<<<ce_grid(n), CE_THREADS, 0, st>>>(n, c, logits, target, ignore_index, acc, loss);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_cross_entropy_forward");
    return CB_OK;
}

extern "C" int cb_cross_entropy_backward(int n, int c, const float *logits, const long long *target, long long ignore_index,
                                         const double *acc, const float *grad_loss, float *grad_logits, void *stream)
{
    CB_REQUIRE(n >= 0 && c > 0 && c <= 1024 && logits && target && acc && grad_loss && grad_logits, CB_EINVAL,
               "cb_cross_entropy_backward: bad arguments");
    if (n == 0) return CB_OK;
    k_ce_backward<<<ce_grid(n), CE_THREADS, 0, (cudaStream_t)stream>>>(n, c, logits, target, ignore_index, acc, grad_loss, grad_logits);
    CB_COUNT(1);
    CB_CUDA_CHECK("cb_cross_entropy_backward");
    return CB_OK;
}
